// Inclusive prefix sum along the middle axis of a row-major [outer, n, inner] view.
//
// Replaces ComputeScan (TensorFrost/Compiler/Implementations.cpp:305-359): one thread per line walking the
// axis serially and storing every prefix.  Here:
//   inner == 1: single-pass chained scan with decoupled look-back.  Each CTA owns one 2048-element
//       tile (256 threads x 8 items, 128-bit loads/stores), publishes its aggregate, and resolves its
//       exclusive prefix by looking back over predecessor tiles of the SAME row.  Tiles are handed out
//       by an atomic ticket so a tile's predecessors are always already running (forward progress).
//       Traffic: n*4 B read + n*4 B written per row — the 2n minimum.
//   inner  > 1: one thread per (outer, inner) column, coalesced along inner, serial along n.
// Integer scans are exact; fp32 scans associate differently from the oracle's serial order (last bits).
#include "tfcuda_internal.h"

namespace {

constexpr int THREADS = 256;
constexpr int ITEMS = 8;
constexpr int TILE = THREADS * ITEMS;

// tile descriptor: high 32 bits = status (0 empty, 1 aggregate ready, 2 inclusive prefix ready), low = value bits
__device__ __forceinline__ unsigned long long pack(unsigned status, unsigned bits) { return ((unsigned long long)status << 32) | bits; }

template <typename T> __device__ __forceinline__ unsigned to_bits(T v);
template <> __device__ __forceinline__ unsigned to_bits<float>(float v) { return __float_as_uint(v); }
template <> __device__ __forceinline__ unsigned to_bits<int>(int v) { return (unsigned)v; }
template <> __device__ __forceinline__ unsigned to_bits<unsigned>(unsigned v) { return v; }
template <typename T> __device__ __forceinline__ T from_bits(unsigned v);
template <> __device__ __forceinline__ float from_bits<float>(unsigned v) { return __uint_as_float(v); }
template <> __device__ __forceinline__ int from_bits<int>(unsigned v) { return (int)v; }
template <> __device__ __forceinline__ unsigned from_bits<unsigned>(unsigned v) { return v; }

template <typename T>
__global__ void __launch_bounds__(THREADS) scan_rows_kernel(const T* __restrict__ in, T* __restrict__ out, size_t n, unsigned tiles_per_row,
                                                            unsigned total_tiles, unsigned long long* __restrict__ desc, unsigned* __restrict__ ticket) {
	__shared__ unsigned s_tile;
	__shared__ T s_warp[THREADS / 32];
	__shared__ T s_prefix;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const unsigned tile = s_tile;
	if (tile >= total_tiles) return;
	const unsigned row = tile / tiles_per_row;
	const unsigned t_in_row = tile - row * tiles_per_row;
	const size_t base = (size_t)t_in_row * TILE;
	const T* p = in + (size_t)row * n + base;
	T* q = out + (size_t)row * n + base;
	const size_t remain = n - base;

	// load 8 consecutive items per thread (blocked arrangement keeps the in-thread scan serial and exact for ints)
	T v[ITEMS];
	const size_t first = (size_t)threadIdx.x * ITEMS;
	const bool vec_ok = ((((size_t)p) & 15) == 0) && (remain >= TILE);
	if (vec_ok) {
		const uint4* p4 = reinterpret_cast<const uint4*>(p + first);
		uint4 a = __ldg(p4), b = __ldg(p4 + 1);
		v[0] = from_bits<T>(a.x); v[1] = from_bits<T>(a.y); v[2] = from_bits<T>(a.z); v[3] = from_bits<T>(a.w);
		v[4] = from_bits<T>(b.x); v[5] = from_bits<T>(b.y); v[6] = from_bits<T>(b.z); v[7] = from_bits<T>(b.w);
	} else {
#pragma unroll
		for (int i = 0; i < ITEMS; i++) v[i] = (first + i < remain) ? p[first + i] : (T)0;
	}
#pragma unroll
	for (int i = 1; i < ITEMS; i++) v[i] = v[i - 1] + v[i];
	T thread_total = v[ITEMS - 1];

	// warp inclusive scan of thread totals
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	T incl = thread_total;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		T o = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl = incl + o;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	T warp_offset = (T)0;
	T block_total = (T)0;
#pragma unroll
	for (int w = 0; w < THREADS / 32; w++) {
		T t = s_warp[w];
		if (w < warp) warp_offset = warp_offset + t;
		block_total = block_total + t;
	}
	T thread_excl = warp_offset + (incl - thread_total);

	// decoupled look-back (first warp)
	if (warp == 0) {
		T prefix = (T)0;
		if (t_in_row == 0) {
			if (lane == 0) atomicExch(&desc[tile], pack(2u, to_bits<T>(block_total)));
		} else {
			if (lane == 0) atomicExch(&desc[tile], pack(1u, to_bits<T>(block_total)));
			int look = (int)tile - 1;
			const int row_first = (int)(row * tiles_per_row);
			for (;;) {
				// each lane inspects one predecessor
				int idx = look - lane;
				unsigned long long d = pack(2u, 0u);  // tiles before the row start count as a finished zero prefix
				if (idx >= row_first) {
					do {
						d = *((volatile unsigned long long*)&desc[idx]);
					} while ((unsigned)(d >> 32) == 0u);
				}
				unsigned status = (unsigned)(d >> 32);
				T val = (idx >= row_first) ? from_bits<T>((unsigned)d) : (T)0;
				unsigned done_mask = __ballot_sync(0xffffffffu, status == 2u);
				// sum the values of lanes up to and including the first finished predecessor
				int stop = done_mask ? (__ffs(done_mask) - 1) : 31;
				T contrib = (lane <= stop) ? val : (T)0;
#pragma unroll
				for (int off = 16; off > 0; off >>= 1) contrib = contrib + __shfl_xor_sync(0xffffffffu, contrib, off);
				prefix = prefix + contrib;
				if (done_mask) break;
				look -= 32;
			}
			if (lane == 0) atomicExch(&desc[tile], pack(2u, to_bits<T>(prefix + block_total)));
		}
		if (lane == 0) s_prefix = prefix;
	}
	__syncthreads();
	const T add = s_prefix + thread_excl;
#pragma unroll
	for (int i = 0; i < ITEMS; i++) v[i] = v[i] + add;
	if (vec_ok) {
		uint4 a, b;
		a.x = to_bits<T>(v[0]); a.y = to_bits<T>(v[1]); a.z = to_bits<T>(v[2]); a.w = to_bits<T>(v[3]);
		b.x = to_bits<T>(v[4]); b.y = to_bits<T>(v[5]); b.z = to_bits<T>(v[6]); b.w = to_bits<T>(v[7]);
		uint4* q4 = reinterpret_cast<uint4*>(q + first);
		q4[0] = a;
		q4[1] = b;
	} else {
#pragma unroll
		for (int i = 0; i < ITEMS; i++)
			if (first + i < remain) q[first + i] = v[i];
	}
}

template <typename T>
__global__ void __launch_bounds__(256) scan_mid_kernel(const T* __restrict__ in, T* __restrict__ out, size_t outer, size_t n, size_t inner) {
	size_t total = outer * inner;
	for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < total; c += (size_t)gridDim.x * blockDim.x) {
		size_t o = c / inner, i = c - o * inner;
		const T* p = in + o * n * inner + i;
		T* q = out + o * n * inner + i;
		T acc = (T)0;
		for (size_t k = 0; k < n; k++) {
			acc = acc + p[k * inner];
			q[k * inner] = acc;
		}
	}
}

template <typename T>
int launch(uint64_t in, uint64_t out, size_t outer, size_t n, size_t inner) {
	tfcuda::State& s = tfcuda::state();
	tfcuda::ProfileScope prof("lib/prefix_sum");
	const T* pin = reinterpret_cast<const T*>(in);
	T* pout = reinterpret_cast<T*>(out);
	if (inner == 1) {
		size_t tiles_per_row = (n + TILE - 1) / TILE;
		size_t total = tiles_per_row * outer;
		if (total > 0x7fffffffull) { tfcuda::set_error("tfcuda_prefix_sum: too many tiles"); return 1; }
		void* scratch = nullptr;
		size_t bytes = total * 8 + 16;
		TFCUDA_CHECK(cudaMallocAsync(&scratch, bytes, s.stream));
		TFCUDA_CHECK(cudaMemsetAsync(scratch, 0, bytes, s.stream));
		unsigned* ticket = reinterpret_cast<unsigned*>(scratch);
		unsigned long long* desc = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(scratch) + 16);
		scan_rows_kernel<T><<<(unsigned)total, THREADS, 0, s.stream>>>(pin, pout, n, (unsigned)tiles_per_row, (unsigned)total, desc, ticket);
		int rc = tfcuda::check_launch("tfcuda_prefix_sum");
		TFCUDA_CHECK(cudaFreeAsync(scratch, s.stream));
		return rc;
	}
	size_t total = outer * inner;
	unsigned blocks = (unsigned)std::min((total + 255) / 256, (size_t)s.sm_count * 16);
	scan_mid_kernel<T><<<blocks ? blocks : 1, 256, 0, s.stream>>>(pin, pout, outer, n, inner);
	return tfcuda::check_launch("tfcuda_prefix_sum");
}

}  // namespace

extern "C" int tfcuda_prefix_sum(uint64_t in, uint64_t out, size_t outer, size_t n, size_t inner, int type) {
	if (!tfcuda::state().initialized) { tfcuda::set_error("tfcuda_prefix_sum: not initialised"); return 1; }
	if (outer == 0 || inner == 0 || n == 0) { tfcuda::set_error("tfcuda_prefix_sum: empty extent"); return 1; }
	switch (type) {
		case TFFloat: return launch<float>(in, out, outer, n, inner);
		case TFInt: return launch<int>(in, out, outer, n, inner);
		case TFUint: return launch<unsigned>(in, out, outer, n, inner);
	}
	tfcuda::set_error("tfcuda_prefix_sum: unsupported element type");
	return 1;
}
