// Store of one 32 x 32 block of an accumulator that arrives in the tcgen05.ld 32x32b layout: lane L of the warp holds ROW L of the
// block, r[j] = column j.  Storing from that layout makes every store instruction of the warp touch 32 different rows (16 bytes in each
// of 32 lines: half-used sectors, 4x the L2 requests of a coalesced store) - and NCA's skinny products are output-dominated
// ([4.2 M, 48] @ [48, 128] reads 0.8 GB and writes 2.1 GB).  Here the 8 float4 units of a lane are transposed with the 8 lanes of its
// octet first (three butterfly stages of __shfl_xor_sync, 48 shuffles), after which lane (octet o, position p) holds columns
// 4p .. 4p+3 of the rows o*8 .. o*8+7, and every store instruction of the warp covers 4 rows x 128 contiguous bytes.
//
// Plain CUDA (no tensor-core state), so it is ALSO compiled for the host and checked there lane by lane against the direct stores
// (tests/cpu_sim/kernel_on_host.cpp, with a barrier-based stand-in for the shuffle).  All 32 lanes must call it (shuffles); rows >= m and
// columns >= n are not written.  `ldc` and `col0` are multiples of 4 and `c` is 16-byte aligned whenever a lane's four columns are all
// inside n (the TMA-describable case this kernel serves), otherwise the lane falls back to scalar stores.
#pragma once

__device__ __forceinline__ void store_block_32x32(uint32_t (&r)[32], float* __restrict__ c, size_t ldc, int row0, int col0, int m, int n, int lane) {
	const int p = lane & 7, o = lane >> 3;
#pragma unroll
	for (int s = 4; s >= 1; s >>= 1) {
		const bool upper = (p & s) != 0;
#pragma unroll
		for (int u = 0; u < 8; u++) {
			if (u & s) continue;  // unit pairs (u, u + s): the lower lane of a pair keeps u and trades u + s, the upper lane the other way round
#pragma unroll
			for (int e = 0; e < 4; e++) {
				const uint32_t send = upper ? r[4 * u + e] : r[4 * (u + s) + e];
				const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, s);
				if (upper) r[4 * u + e] = recv;
				else r[4 * (u + s) + e] = recv;
			}
		}
	}
	const int col = col0 + 4 * p;
#pragma unroll
	for (int u = 0; u < 8; u++) {
		const int row = row0 + o * 8 + u;
		if (row >= m) continue;
		float* dst = c + (size_t)row * ldc + col;
		if (col + 4 <= n) {
			*reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]), __uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 3]));
		} else {
#pragma unroll
			for (int e = 0; e < 4; e++)
				if (col + e < n) dst[e] = __uint_as_float(r[4 * u + e]);
		}
	}
}
