// C (R x N) = A (R x K) @ B (K x N), row-major fp32, for a SMALL weight matrix B applied to very many rows
// (NCA: R = B*H*W = 4.2 M rows against 48x128, 128x12, 12x128 and 128x48 weights).
//
// EXPERIMENTAL in round 1: written after the round's GPU budget ended, compiled for sm_100a but never run on hardware.
// It is reachable only through tfcuda_matmul_rows (tests/test_library_gpu.py::test_matmul_rows, skipped unless
// TFCUDA_EXPERIMENTAL=1) and, inside compiled programs, when TFCUDA_MATMUL_ROWS=1; the default path is unchanged.
//
// Why: these products are HBM-bound (fc1 forward reads 0.8 GB and writes 2.1 GB for 51 GFLOP), and the generic tcgen05 path
// (matmul_tcgen05.cu) pays a hi/lo split pre-pass over A (12 B moved per element of A) plus a transposed copy of B per call:
// 1.85 ms per product where the traffic alone needs 0.45 ms.  Here B (at most 96 KB) is staged in shared memory ONCE per CTA,
// CTAs are persistent over row tiles, A streams through a double-buffered shared-memory chunk exactly as it lies in memory
// (a tile of whole rows is one contiguous block: coalesced 128-bit loads, no transpose), and the products are plain fp32 FFMA in
// the reference's k order (Compiler/Implementations.cpp:560-646) - bit-compatible with the oracle up to FMA contraction, no TF32.
// Bound: fp32 FFMA issue (2*K*N flop per row against 4*(K+N) bytes).
#ifndef TF_HOST_SIM  // tests/cpu_sim/kernel_on_host.cpp compiles the kernel below for the host through cuda_host_shim.h
#include <algorithm>

#include "tfcuda_internal.h"
#endif

namespace {

constexpr int MR_THREADS = 256;
constexpr int MR_KC = 32;       // k-extent of one A chunk in shared memory
constexpr int MR_APITCH = 36;   // floats per A row in shared memory (KC + 4: keeps rows 16-byte aligned, staggers banks)

// TXN threads along n, each owning CH float4 column chunks (chunk c covers columns c*(BN/CH) + tx*4 .. +3, so that the 8 lanes of
// a quarter-warp read 128 contiguous bytes of a B row); TM rows per thread.  BN = TXN*4*CH columns, BR = (256/TXN)*TM rows per tile.
template <int TXN, int CH, int TM>
__global__ void __launch_bounds__(MR_THREADS) matmul_rows_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                                                               long long R, int K, int N, long long tiles) {
	constexpr int BN = TXN * 4 * CH;
	constexpr int TY = MR_THREADS / TXN;
	constexpr int BR = TY * TM;
	extern __shared__ __align__(16) float smem[];
	const int kpad = (K + 3) & ~3;
	float* Bs = smem;                                  // [kpad][BN], columns >= N and rows >= K are zero
	float* As = smem + (size_t)kpad * BN;               // [2][BR][MR_APITCH]
	const int tid = threadIdx.x;
	const int tx = tid % TXN, ty = tid / TXN;

	for (int e = tid; e < kpad * BN; e += MR_THREADS) {
		const int k = e / BN, n = e - k * BN;
		Bs[e] = (k < K && n < N) ? __ldg(B + (size_t)k * N + n) : 0.0f;
	}

	constexpr int VECS = BR * (MR_KC / 4);                       // float4 slots of one A chunk
	constexpr int PER = (VECS + MR_THREADS - 1) / MR_THREADS;
	const int chunks = (K + MR_KC - 1) / MR_KC;
	const bool vec_ok = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);

	for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		const long long r0 = tile * BR;
		float acc[TM][4 * CH];
#pragma unroll
		for (int i = 0; i < TM; i++)
#pragma unroll
			for (int j = 0; j < 4 * CH; j++) acc[i][j] = 0.0f;
		float4 stage[PER];

		auto fetch = [&](int chunk) {
			const int k0 = chunk * MR_KC;
#pragma unroll
			for (int q = 0; q < PER; q++) {
				const int v = tid + q * MR_THREADS;
				const int rr = v / (MR_KC / 4), kq = (v % (MR_KC / 4)) * 4;
				const long long gr = r0 + rr;
				const int gk = k0 + kq;
				float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
				if (v < VECS && gr < R && gk < K) {
					const float* p = A + gr * K + gk;
					if (vec_ok) {  // K % 4 == 0: the four words never straddle a row end
						val = __ldg(reinterpret_cast<const float4*>(p));
					} else {
						val.x = __ldg(p);
						if (gk + 1 < K) val.y = __ldg(p + 1);
						if (gk + 2 < K) val.z = __ldg(p + 2);
						if (gk + 3 < K) val.w = __ldg(p + 3);
					}
				}
				stage[q] = val;
			}
		};
		auto stash = [&](int buf) {
#pragma unroll
			for (int q = 0; q < PER; q++) {
				const int v = tid + q * MR_THREADS;
				if (v < VECS) {
					const int rr = v / (MR_KC / 4), kq = (v % (MR_KC / 4)) * 4;
					*reinterpret_cast<float4*>(&As[((size_t)buf * BR + rr) * MR_APITCH + kq]) = stage[q];
				}
			}
		};

		fetch(0);
		__syncthreads();  // Bs complete (first tile) / every thread is done with both A buffers of the previous tile
		stash(0);
		__syncthreads();
		int buf = 0;
		for (int chunk = 0; chunk < chunks; chunk++) {
			const bool more = chunk + 1 < chunks;
			if (more) fetch(chunk + 1);
			const int k0 = chunk * MR_KC;
			const int klen = min(MR_KC, kpad - k0);  // multiple of 4; A words beyond K are zero, so are the B rows beyond K
			for (int k = 0; k < klen; k += 4) {
				float4 a4[TM];
#pragma unroll
				for (int i = 0; i < TM; i++) a4[i] = *reinterpret_cast<const float4*>(&As[((size_t)buf * BR + ty * TM + i) * MR_APITCH + k]);
#pragma unroll
				for (int kk = 0; kk < 4; kk++) {
					const float* brow = Bs + (size_t)(k0 + k + kk) * BN;
#pragma unroll
					for (int c = 0; c < CH; c++) {
						const float4 b4 = *reinterpret_cast<const float4*>(brow + c * (BN / CH) + tx * 4);
#pragma unroll
						for (int i = 0; i < TM; i++) {
							const float a = kk == 0 ? a4[i].x : kk == 1 ? a4[i].y : kk == 2 ? a4[i].z : a4[i].w;
							acc[i][c * 4 + 0] = fmaf(a, b4.x, acc[i][c * 4 + 0]);
							acc[i][c * 4 + 1] = fmaf(a, b4.y, acc[i][c * 4 + 1]);
							acc[i][c * 4 + 2] = fmaf(a, b4.z, acc[i][c * 4 + 2]);
							acc[i][c * 4 + 3] = fmaf(a, b4.w, acc[i][c * 4 + 3]);
						}
					}
				}
			}
			if (more) stash(buf ^ 1);
			__syncthreads();
			buf ^= 1;
		}

		const bool cvec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
		for (int i = 0; i < TM; i++) {
			const long long gr = r0 + ty * TM + i;
			if (gr >= R) continue;
			float* crow = C + gr * N;
#pragma unroll
			for (int c = 0; c < CH; c++) {
				const int col = c * (BN / CH) + tx * 4;
				if (cvec && col + 4 <= N) {
					*reinterpret_cast<float4*>(crow + col) = make_float4(acc[i][c * 4 + 0], acc[i][c * 4 + 1], acc[i][c * 4 + 2], acc[i][c * 4 + 3]);
				} else {
#pragma unroll
					for (int j = 0; j < 4; j++)
						if (col + j < N) crow[col + j] = acc[i][c * 4 + j];
				}
			}
		}
	}
}

#ifndef TF_HOST_SIM
template <int TXN, int CH, int TM>
int launch_rows(const float* a, const float* b, float* c, size_t r, size_t k, size_t n) {
	tfcuda::State& s = tfcuda::state();
	constexpr int BN = TXN * 4 * CH;
	constexpr int BR = (MR_THREADS / TXN) * TM;
	const size_t kpad = (k + 3) & ~size_t(3);
	const size_t smem_bytes = (kpad * BN + 2 * (size_t)BR * MR_APITCH) * sizeof(float);
	if (smem_bytes > 200 * 1024) {
		tfcuda::set_error("tfcuda_matmul_rows: the weight matrix does not fit in shared memory");
		return 1;
	}
	static size_t configured = 0;  // per template instantiation
	if (smem_bytes > configured) {
		TFCUDA_CHECK(cudaFuncSetAttribute(matmul_rows_kernel<TXN, CH, TM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
		configured = smem_bytes;
	}
	const long long tiles = (long long)((r + BR - 1) / BR);
	const size_t per_sm = std::max<size_t>(1, std::min<size_t>(4, (220 * 1024) / smem_bytes));
	const unsigned grid = (unsigned)std::min<long long>(tiles, (long long)s.sm_count * (long long)per_sm);
	tfcuda::ProfileScope prof("lib/matmul_rows", 4.0 * (double)r * (double)(k + n));
	matmul_rows_kernel<TXN, CH, TM><<<grid, MR_THREADS, smem_bytes, s.stream>>>(a, b, c, (long long)r, (int)k, (int)n, tiles);
	return tfcuda::check_launch("matmul_rows_kernel");
}

#endif  // TF_HOST_SIM

}  // namespace

#ifndef TF_HOST_SIM
extern "C" int tfcuda_matmul_rows_supported(size_t r, size_t k, size_t n) {
	if (r == 0 || k == 0 || n == 0 || n > 128 || k > 0x7fff) return 0;
	const size_t bn = n <= 16 ? 16 : n <= 32 ? 32 : n <= 64 ? 64 : 128;
	const size_t br = n <= 16 ? 256 : n <= 32 ? 128 : 64;
	return (((k + 3) & ~size_t(3)) * bn + 2 * br * MR_APITCH) * sizeof(float) <= 200 * 1024;
}

extern "C" int tfcuda_matmul_rows(uint64_t a, uint64_t b, uint64_t c, size_t r, size_t k, size_t n) {
	tfcuda::State& s = tfcuda::state();
	if (!s.initialized) { tfcuda::set_error("tfcuda_matmul_rows: not initialised"); return 1; }
	if (r == 0 || n == 0) return 0;
	if (k == 0) return tfcuda_memset32(c, 0, r * n);
	if (!tfcuda_matmul_rows_supported(r, k, n)) { tfcuda::set_error("tfcuda_matmul_rows: needs n <= 128 and a weight matrix that fits in shared memory"); return 1; }
	const float* pa = reinterpret_cast<const float*>(a);
	const float* pb = reinterpret_cast<const float*>(b);
	float* pc = reinterpret_cast<float*>(c);
	if (n <= 16) return launch_rows<4, 1, 4>(pa, pb, pc, r, k, n);    // BN 16,  BR 256
	if (n <= 32) return launch_rows<8, 1, 4>(pa, pb, pc, r, k, n);    // BN 32,  BR 128
	if (n <= 64) return launch_rows<16, 1, 4>(pa, pb, pc, r, k, n);   // BN 64,  BR 64
	return launch_rows<16, 2, 4>(pa, pb, pc, r, k, n);                // BN 128, BR 64
}
#endif  // TF_HOST_SIM
