// dst[index[i]] += src[i]   — tf.scatterAdd and the autodiff of `load` (Implementations.cpp:185-194).
//
// The reference emulates float atomic add with a compare-and-swap loop (CPP.cpp:150-158, GLSL.cpp:121-134;
// "extremely slow with high write contention", README.md:825).  Here the add is the native
// red.global.add.{f32,s32,u32}, and lanes of a warp that hit the SAME address are first combined in
// registers (__match_any_sync groups, the group's lowest lane issues one atomic), so a warp issues one
// atomic per DISTINCT address.  Indices are clamped to the destination like the reference's default
// indexing mode (Steps/GraphOps.cpp:1022-1025).  Traffic: 8 B read per element + the atomics.
#include "tfcuda_internal.h"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) scatter_add_kernel(T* __restrict__ dst, const int* __restrict__ index, const T* __restrict__ src, size_t n, int last) {
	const int lane = threadIdx.x & 31;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	// whole warps iterate together so the match/shuffle masks are always full
	const size_t n_round = (n + 31) & ~size_t(31);
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
		const bool valid = i < n;
		int idx = valid ? index[i] : -1 - lane;  // invalid lanes get unique keys, never written
		T val = valid ? src[i] : (T)0;
		if (valid) idx = min(max(idx, 0), last);
		unsigned peers = __match_any_sync(0xffffffffu, idx);
		int members = __popc(peers);
		int max_members = __reduce_max_sync(0xffffffffu, members);
		if (max_members > 1) {
			// combine the group in lane order (deterministic inside the warp)
			T sum = (T)0;
			for (int k = 0; k < max_members; k++) {
				int from = k < members ? (int)__fns(peers, 0, k + 1) : lane;
				T other = __shfl_sync(0xffffffffu, val, from);
				if (k < members) sum += other;
			}
			val = sum;
		}
		if (valid && lane == __ffs(peers) - 1) atomicAdd(dst + idx, val);
	}
}

template <typename T>
int launch(uint64_t dst, uint64_t index, uint64_t src, size_t n, size_t dst_words) {
	tfcuda::State& s = tfcuda::state();
	tfcuda::ProfileScope prof("lib/scatter_add");
	unsigned blocks = (unsigned)std::min((n + 255) / 256, (size_t)s.sm_count * 16);
	scatter_add_kernel<T><<<blocks ? blocks : 1, 256, 0, s.stream>>>(reinterpret_cast<T*>(dst), reinterpret_cast<const int*>(index),
	                                                                   reinterpret_cast<const T*>(src), n, (int)dst_words - 1);
	return tfcuda::check_launch("tfcuda_scatter_add");
}

}  // namespace

extern "C" int tfcuda_scatter_add(uint64_t dst, uint64_t index, uint64_t src, size_t n, size_t dst_words, int type) {
	if (!tfcuda::state().initialized) { tfcuda::set_error("tfcuda_scatter_add: not initialised"); return 1; }
	if (n == 0) return 0;
	if (dst_words == 0 || dst_words > 0x7fffffffull) { tfcuda::set_error("tfcuda_scatter_add: destination size out of range"); return 1; }
	switch (type) {
		case TFFloat: return launch<float>(dst, index, src, n, dst_words);
		case TFInt: return launch<int>(dst, index, src, n, dst_words);
		case TFUint: return launch<unsigned>(dst, index, src, n, dst_words);
	}
	tfcuda::set_error("tfcuda_scatter_add: unsupported element type");
	return 1;
}
