// Stable LSD radix sort of 32-bit keys (+ optional 32-bit values): "one-sweep" with 8-bit digits.
//
// Replaces the user-level sort of the reference (Python/TensorFrost/sort.py:38-187): 6-bit digits,
// 6 passes, and per pass a packed histogram kernel, a two-level grouped scan and a scatter kernel whose
// threads each re-count their group serially (13 kernels, 37 dispatches per sort).  Same contract:
// stable, ascending, low `max_bits` bits of the mapped key, with the key bijections of sort.py:52-72
// (float: IEEE total order, int: sign flip).  The result is bit-identical to the reference's
// (stable sort == np.argsort(kind="stable")).
//
// Algorithm (Adinets & Merrill's Onesweep shape, written from scratch for sm_100a):
//   1. digit_histogram_kernel: ONE read of the keys builds the histograms of all passes (shared-memory
//      atomics, then one global atomic per bin), 2. scan_histogram_kernel turns them into exclusive
//      digit offsets, 3. per pass, onesweep_kernel: each CTA takes an 8192-key tile by atomic ticket,
//      ranks its keys stably with a ballot-per-bit warp multi-split, resolves its per-digit tile offset
//      by decoupled look-back (256 digits = 256 threads looking back in parallel), stages the tile
//      digit-sorted in shared memory and writes runs of equal digit to consecutive global addresses.
//      (Round 2 tried a warp-per-digit look-back over a digit-major descriptor table, 32 predecessors per step: 12.75 ms instead of
//      10.44 ms for 2^28 keys on the B200 - the serial walk is short in practice and the extra barriers cost more; reverted.)
// Traffic per pass: n*4 B read + n*4 B written for keys (same again for values) — the minimum for an
// out-of-place pass — plus the 1 KB/tile look-back descriptors (3%).  Keys-only, 4 passes:
// (1 + 2*4)*4 = 36 B/key.
#include "tfcuda_internal.h"

namespace {

constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 8192 keys per CTA
constexpr int RS_BINS = 256;

constexpr unsigned FLAG_AGG = 1u << 30;
constexpr unsigned FLAG_PREFIX = 2u << 30;
constexpr unsigned VALUE_MASK = (1u << 30) - 1;

// key bijections of sort.py:52-72 (mode: 0 none, 1 int, 2 float)
__device__ __forceinline__ unsigned key_to_bits(unsigned k, int mode) {
	if (mode == 1) return k ^ 0x80000000u;
	if (mode == 2) return k ^ ((k >> 31) ? 0xffffffffu : 0x80000000u);
	return k;
}
__device__ __forceinline__ unsigned bits_to_key(unsigned k, int mode) {
	if (mode == 1) return k ^ 0x80000000u;
	if (mode == 2) return k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu);
	return k;
}

// ---- 1. histograms of every pass in one read -----------------------------------------------------
__global__ void __launch_bounds__(512) digit_histogram_kernel(const unsigned* __restrict__ keys, size_t n, int mode, int passes, int max_bits,
                                                              unsigned* __restrict__ hist /* [passes][256] */) {
	__shared__ unsigned s_hist[4][RS_BINS];
	for (int i = threadIdx.x; i < 4 * RS_BINS; i += blockDim.x) (&s_hist[0][0])[i] = 0;
	__syncthreads();
	const size_t n4 = ((reinterpret_cast<size_t>(keys) & 15) == 0) ? (n >> 2) : 0;  // 128-bit loads need alignment
	const uint4* keys4 = reinterpret_cast<const uint4*>(keys);
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	auto count = [&](unsigned raw) {
		unsigned k = key_to_bits(raw, mode);
#pragma unroll
		for (int p = 0; p < 4; p++) {
			if (p < passes) {
				int bits = min(8, max_bits - 8 * p);
				unsigned d = (k >> (8 * p)) & ((1u << bits) - 1u);
				atomicAdd(&s_hist[p][d], 1u);
			}
		}
	};
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
		uint4 w = __ldg(keys4 + i);
		count(w.x); count(w.y); count(w.z); count(w.w);
	}
	for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) count(keys[i]);
	__syncthreads();
	for (int i = threadIdx.x; i < passes * RS_BINS; i += blockDim.x) {
		unsigned c = (&s_hist[0][0])[i];
		if (c) atomicAdd(&hist[i], c);
	}
}

// ---- 2. exclusive scan of each 256-bin histogram (one warp-synchronous block per pass) -------------
__global__ void __launch_bounds__(RS_BINS) scan_histogram_kernel(unsigned* __restrict__ hist) {
	__shared__ unsigned s_warp[RS_BINS / 32];
	unsigned* h = hist + blockIdx.x * RS_BINS;
	unsigned v = h[threadIdx.x];
	unsigned incl = v;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += o;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	unsigned base = 0;
	for (int w = 0; w < warp; w++) base += s_warp[w];
	h[threadIdx.x] = base + incl - v;
}

// ---- 3. one pass ------------------------------------------------------------------------------------
struct PassArgs {
	const unsigned* keys_in;
	unsigned* keys_out;
	const unsigned* vals_in;
	unsigned* vals_out;
	size_t n;
	int shift;
	unsigned digit_mask;
	int mode_in;   // key bijection applied when loading (first pass only)
	int mode_out;  // inverse bijection applied when storing (last pass only)
	const unsigned* digit_base;  // [256] exclusive global offsets of this pass
	unsigned* desc;              // [tiles][256] look-back descriptors, zero-initialised
	unsigned* ticket;            // zero-initialised
};

template <bool HAS_VALUES>
__global__ void __launch_bounds__(RS_THREADS, 2) onesweep_kernel(const __grid_constant__ PassArgs a) {
	extern __shared__ unsigned smem[];
	unsigned* s_keys = smem;                                   // [RS_TILE]
	unsigned* s_warp_hist = s_keys + RS_TILE;                  // [RS_WARPS][256]
	unsigned* s_digit_base = s_warp_hist + RS_WARPS * RS_BINS; // [256] start of each digit inside the sorted tile
	unsigned* s_global_off = s_digit_base + RS_BINS;           // [256] global address of sorted position 0 of the digit, minus s_digit_base
	unsigned* s_scan = s_global_off + RS_BINS;                 // [8]
	unsigned* s_vals = s_scan + 32;                            // [RS_TILE] when HAS_VALUES
	__shared__ unsigned s_tile;

	if (threadIdx.x == 0) s_tile = atomicAdd(a.ticket, 1u);
	for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) s_warp_hist[i] = 0;
	__syncthreads();
	const unsigned tile = s_tile;
	const size_t tile_base = (size_t)tile * RS_TILE;
	const unsigned valid_in_tile = (unsigned)min((size_t)RS_TILE, a.n - tile_base);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const unsigned lane_lt = (1u << lane) - 1u;

	// warp-striped load: item i of lane l is element warp*512 + i*32 + l of the tile (coalesced, order preserving)
	unsigned key[RS_ITEMS];
	unsigned short rank[RS_ITEMS];
	const unsigned warp_first = warp * (32 * RS_ITEMS) + lane;
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		unsigned e = warp_first + i * 32;
		key[i] = e < valid_in_tile ? key_to_bits(a.keys_in[tile_base + e], a.mode_in) : 0xffffffffu;
	}

	// stable multi-split inside the warp: keys with equal digit form a match group; the group's lowest lane
	// advances the warp's digit counter, every member takes counter + (#members in lower lanes)
	unsigned* my_hist = s_warp_hist + warp * RS_BINS;
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		unsigned e = warp_first + i * 32;
		bool valid = e < valid_in_tile;
		unsigned d = valid ? ((key[i] >> a.shift) & a.digit_mask) : RS_BINS;  // invalid lanes match only each other
		// lanes with the same digit: one ballot per digit bit (+ validity) instead of match.any - ncu (round 2, 2^26 keys): with
		// MATCH.ANY the kernel sat at 72 % of its busiest pipe with 21 % issue activity and 10 % of DRAM; the vote pipe is far cheaper
		unsigned peers = __ballot_sync(0xffffffffu, valid);
		peers = valid ? peers : ~peers;
#pragma unroll
		for (int b = 0; b < 8; b++) {
			const bool bit = (d >> b) & 1u;
			const unsigned m = __ballot_sync(0xffffffffu, bit);
			peers &= bit ? m : ~m;
		}
		int leader = __ffs(peers) - 1;
		unsigned before = 0;
		if (lane == leader && valid) {
			before = my_hist[d];
			my_hist[d] = before + __popc(peers);
		}
		before = __shfl_sync(0xffffffffu, before, leader);
		rank[i] = (unsigned short)(before + __popc(peers & lane_lt));
		__syncwarp();
	}
	__syncthreads();

	// per digit: exclusive offsets over warps, tile total, look-back, tile-local digit base
	unsigned tile_count = 0;
	if (threadIdx.x < RS_BINS) {
		const int d = threadIdx.x;
#pragma unroll
		for (int w = 0; w < RS_WARPS; w++) {
			unsigned c = s_warp_hist[w * RS_BINS + d];
			s_warp_hist[w * RS_BINS + d] = tile_count;
			tile_count += c;
		}
		// publish the aggregate as early as possible
		unsigned* my_desc = a.desc + (size_t)tile * RS_BINS + d;
		if (tile == 0) atomicExch(my_desc, FLAG_PREFIX | tile_count);
		else atomicExch(my_desc, FLAG_AGG | tile_count);
		// exclusive scan of tile_count over the 256 digits
		unsigned incl = tile_count;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			unsigned o = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= off) incl += o;
		}
		if (lane == 31) s_scan[warp] = incl;
		// named barrier over the first 8 warps only
		asm volatile("bar.sync 1, 256;");
		unsigned base = 0;
		for (int w = 0; w < warp; w++) base += s_scan[w];
		unsigned digit_start = base + incl - tile_count;
		s_digit_base[d] = digit_start;
		// decoupled look-back over predecessor tiles for this digit
		unsigned excl = 0;
		if (tile > 0) {
			for (long t = (long)tile - 1; t >= 0; t--) {
				const volatile unsigned* p = a.desc + (size_t)t * RS_BINS + d;
				unsigned v;
				do { v = *p; } while ((v & ~VALUE_MASK) == 0u);
				excl += v & VALUE_MASK;
				if ((v & ~VALUE_MASK) == FLAG_PREFIX) break;
			}
			atomicExch(my_desc, FLAG_PREFIX | (excl + tile_count));
		}
		s_global_off[d] = a.digit_base[d] + excl - digit_start;
	}
	__syncthreads();

	// scatter into the digit-sorted tile in shared memory
	unsigned short pos[RS_ITEMS];
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		unsigned e = warp_first + i * 32;
		if (e < valid_in_tile) {
			unsigned d = (key[i] >> a.shift) & a.digit_mask;
			unsigned p = s_digit_base[d] + my_hist[d] + rank[i];
			pos[i] = (unsigned short)p;
			s_keys[p] = key[i];
		}
	}
	if (HAS_VALUES) {
#pragma unroll
		for (int i = 0; i < RS_ITEMS; i++) {
			unsigned e = warp_first + i * 32;
			if (e < valid_in_tile) s_vals[pos[i]] = a.vals_in[tile_base + e];
		}
	}
	__syncthreads();

	// write out: consecutive sorted positions of one digit go to consecutive global addresses
#pragma unroll 4
	for (unsigned j = threadIdx.x; j < valid_in_tile; j += RS_THREADS) {
		unsigned k = s_keys[j];
		unsigned d = (k >> a.shift) & a.digit_mask;
		unsigned dst = s_global_off[d] + j;
		a.keys_out[dst] = bits_to_key(k, a.mode_out);
		if (HAS_VALUES) a.vals_out[dst] = s_vals[j];
	}
}

constexpr size_t smem_bytes(bool has_values) {
	return (size_t)(RS_TILE + RS_WARPS * RS_BINS + RS_BINS + RS_BINS + 32 + (has_values ? RS_TILE : 0)) * 4;
}

size_t tiles_of(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

}  // namespace

extern "C" size_t tfcuda_radix_sort_temp_words(size_t n) {
	// alternate key buffer + alternate value buffer + 4 histograms + 4 tickets (padded) + 4 descriptor tables
	return 2 * n + 4 * RS_BINS + 64 + 4 * tiles_of(n) * RS_BINS + 64;
}

extern "C" int tfcuda_radix_sort(uint64_t keys_in, uint64_t keys_out, uint64_t values_in, uint64_t values_out, size_t n, int key_type,
                                 int max_bits, uint64_t temp) {
	tfcuda::State& s = tfcuda::state();
	if (!s.initialized) { tfcuda::set_error("tfcuda_radix_sort: not initialised"); return 1; }
	if (n == 0) return 0;
	if (n >= (size_t)VALUE_MASK) { tfcuda::set_error("tfcuda_radix_sort: at most 2^30-1 keys"); return 1; }
	if (max_bits < 1 || max_bits > 32) { tfcuda::set_error("tfcuda_radix_sort: max_bits must be in [1,32]"); return 1; }
	if (!keys_in || !keys_out || !temp) { tfcuda::set_error("tfcuda_radix_sort: null buffer"); return 1; }
	const bool has_values = values_in != 0;
	if (has_values && !values_out) { tfcuda::set_error("tfcuda_radix_sort: values_out missing"); return 1; }
	const int mode = key_type == TFFloat ? 2 : (key_type == TFInt ? 1 : 0);
	const int passes = (max_bits + 7) / 8;
	const size_t tiles = tiles_of(n);

	unsigned* t = reinterpret_cast<unsigned*>(temp);
	unsigned* alt_keys = t;
	unsigned* alt_vals = t + n;
	unsigned* meta = t + 2 * n;
	meta = reinterpret_cast<unsigned*>((reinterpret_cast<uintptr_t>(meta) + 127) & ~uintptr_t(127));
	unsigned* hist = meta;                 // [4][256]
	unsigned* tickets = hist + 4 * RS_BINS;  // [4] (padded to 32)
	unsigned* desc = tickets + 32;         // [4][tiles][256]
	size_t meta_words = 4 * RS_BINS + 32 + (size_t)passes * tiles * RS_BINS;
	TFCUDA_CHECK(cudaMemsetAsync(meta, 0, meta_words * 4, s.stream));

	static bool attr_set = false;
	if (!attr_set) {
		TFCUDA_CHECK(cudaFuncSetAttribute(onesweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(false)));
		TFCUDA_CHECK(cudaFuncSetAttribute(onesweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(true)));
		attr_set = true;
	}

	const unsigned* kin = reinterpret_cast<const unsigned*>(keys_in);
	unsigned hist_blocks = (unsigned)std::min<size_t>((n / 4 + 511) / 512 + 1, (size_t)s.sm_count * 4);
	{
		tfcuda::ProfileScope prof("lib/radix_histogram", 4.0 * n);
		digit_histogram_kernel<<<hist_blocks, 512, 0, s.stream>>>(kin, n, mode, passes, max_bits, hist);
		if (tfcuda::check_launch("digit_histogram_kernel")) return 1;
	}
	scan_histogram_kernel<<<passes, RS_BINS, 0, s.stream>>>(hist);
	if (tfcuda::check_launch("scan_histogram_kernel")) return 1;

	const unsigned* src_k = kin;
	const unsigned* src_v = reinterpret_cast<const unsigned*>(values_in);
	for (int p = 0; p < passes; p++) {
		bool to_out = ((passes - 1 - p) & 1) == 0;
		PassArgs a;
		a.keys_in = src_k;
		a.keys_out = to_out ? reinterpret_cast<unsigned*>(keys_out) : alt_keys;
		a.vals_in = src_v;
		a.vals_out = has_values ? (to_out ? reinterpret_cast<unsigned*>(values_out) : alt_vals) : nullptr;
		a.n = n;
		a.shift = 8 * p;
		int bits = std::min(8, max_bits - 8 * p);
		a.digit_mask = (1u << bits) - 1u;
		a.mode_in = p == 0 ? mode : 0;
		a.mode_out = p == passes - 1 ? mode : 0;
		a.digit_base = hist + p * RS_BINS;
		a.desc = desc + (size_t)p * tiles * RS_BINS;
		a.ticket = tickets + p;
		tfcuda::ProfileScope prof("lib/radix_onesweep", (has_values ? 16.0 : 8.0) * n);
		if (has_values) onesweep_kernel<true><<<(unsigned)tiles, RS_THREADS, smem_bytes(true), s.stream>>>(a);
		else onesweep_kernel<false><<<(unsigned)tiles, RS_THREADS, smem_bytes(false), s.stream>>>(a);
		if (tfcuda::check_launch("onesweep_kernel")) return 1;
		src_k = a.keys_out;
		src_v = a.vals_out;
	}
	return 0;
}
