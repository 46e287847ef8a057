// C (M x N) = A^T @ B with A [R x M] and B [R x N], row-major fp32, contracted over the LEADING extent R.
//
// This is the weight-gradient contraction of `Y = X @ W`:  dW = X^T @ dY with R = every row of the batch (NCA: R = B*H*W =
// 4.2 M, M x N = 48 x 128 and 128 x 12).  The reference's VJP (Compiler/Implementations.cpp:133-135) writes it as a batched
// matmul Transpose(X)[b] @ dY[b] per leading slice, materialises the [batch, M, N] products (805 MB per CA step at the NCA
// config) and then sums them over the batch axes with separate reduction kernels (ComputeMatMul :560-646 + ComputeReduction
// :243-303).  Here the whole thing is ONE pass over X and dY:
//   * both operands are read exactly as they lie in memory (row r of A and of B are contiguous): coalesced 128-bit loads, no
//     transposed copy; a CTA stages BR rows of each into shared memory and every thread keeps a TM x TN register tile of C;
//   * the contraction is split over gridDim.y CTAs (split-K, R is 10^6 while M*N is one or two tiles), each writes its partial
//     tile to a workspace, and a second kernel adds the partials in a fixed order: deterministic, no float atomics.
// Bound: fp32 FFMA issue (2*R*M*N flop against 4*R*(M+N) bytes is ~18 flop/byte at 48 x 128, above the FFMA/HBM balance point).
#ifndef TF_HOST_SIM  // tests/cpu_sim/kernel_on_host.cpp compiles the kernels below for the host through cuda_host_shim.h
#include <algorithm>

#include "tfcuda_internal.h"
#endif

namespace {

constexpr int TN_BR = 16;  // rows of A / B per shared-memory stage

template <int BM, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) matmul_tn_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ partial,
                                                                          long long R, int M, int N, long long rows_per_split, int tiles_n) {
	constexpr int THREADS = (BM / TM) * (BN / TN);
	__shared__ __align__(16) float As[2][TN_BR][BM];
	__shared__ __align__(16) float Bs[2][TN_BR][BN];
	const int tile = blockIdx.x;
	const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
	const long long r_begin = (long long)blockIdx.y * rows_per_split;
	const long long r_end = min(R, r_begin + rows_per_split);
	const int tid = threadIdx.x;
	const int tx = tid % (BN / TN), ty = tid / (BN / TN);
	// Columns of a thread's register tile.  A thread whose TN columns were contiguous (tx * TN ...) would read shared memory at a stride of
	// TN words across the lanes of a warp: with TN = 8 four lanes share a bank on every b load (4-way conflict, and the b loads outnumber
	// the a loads 2:1).  So a TN that is a multiple of 4 is split into groups of 4 columns, group g at g * GROUP_STRIDE + tx * 4: the
	// lanes of a warp then read consecutive 16-byte words (one conflict-free LDS.128 per group), and global stores of a row of C are
	// 16 bytes per lane, consecutive across lanes.  Narrow tiles (TN = 2) keep the contiguous layout (8-byte stride: conflict-free).
	constexpr bool GROUPED = (TN % 4 == 0);
	constexpr int GROUP_STRIDE = (BN / TN) * 4;
	auto column = [&](int j) { return GROUPED ? (j / 4) * GROUP_STRIDE + tx * 4 + (j % 4) : tx * TN + j; };

	float acc[TM][TN];
#pragma unroll
	for (int i = 0; i < TM; i++)
#pragma unroll
		for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;

	// global -> registers -> shared, one stage ahead of the math
	constexpr int A_VECS = TN_BR * BM / 4, B_VECS = TN_BR * BN / 4;
	constexpr int A_PER = (A_VECS + THREADS - 1) / THREADS, B_PER = (B_VECS + THREADS - 1) / THREADS;
	float4 ra[A_PER], rb[B_PER];

	auto fetch = [&](long long r0) {
#pragma unroll
		for (int q = 0; q < A_PER; q++) {
			const int v = tid + q * THREADS;
			const int rr = v / (BM / 4), cc = (v % (BM / 4)) * 4;
			const long long gr = r0 + rr;
			const int gc = m0 + cc;
			float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
			if (v < A_VECS && gr < r_end) {
				const float* p = A + gr * M + gc;
				if (VEC && gc + 4 <= M) {
					val = __ldg(reinterpret_cast<const float4*>(p));
				} else {
					if (gc + 0 < M) val.x = __ldg(p + 0);
					if (gc + 1 < M) val.y = __ldg(p + 1);
					if (gc + 2 < M) val.z = __ldg(p + 2);
					if (gc + 3 < M) val.w = __ldg(p + 3);
				}
			}
			ra[q] = val;
		}
#pragma unroll
		for (int q = 0; q < B_PER; q++) {
			const int v = tid + q * THREADS;
			const int rr = v / (BN / 4), cc = (v % (BN / 4)) * 4;
			const long long gr = r0 + rr;
			const int gc = n0 + cc;
			float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
			if (v < B_VECS && gr < r_end) {
				const float* p = B + gr * N + gc;
				if (VEC && gc + 4 <= N) {
					val = __ldg(reinterpret_cast<const float4*>(p));
				} else {
					if (gc + 0 < N) val.x = __ldg(p + 0);
					if (gc + 1 < N) val.y = __ldg(p + 1);
					if (gc + 2 < N) val.z = __ldg(p + 2);
					if (gc + 3 < N) val.w = __ldg(p + 3);
				}
			}
			rb[q] = val;
		}
	};
	auto stash = [&](int buf) {
#pragma unroll
		for (int q = 0; q < A_PER; q++) {
			const int v = tid + q * THREADS;
			if (v < A_VECS) *reinterpret_cast<float4*>(&As[buf][v / (BM / 4)][(v % (BM / 4)) * 4]) = ra[q];
		}
#pragma unroll
		for (int q = 0; q < B_PER; q++) {
			const int v = tid + q * THREADS;
			if (v < B_VECS) *reinterpret_cast<float4*>(&Bs[buf][v / (BN / 4)][(v % (BN / 4)) * 4]) = rb[q];
		}
	};

	if (r_begin < r_end) {
		fetch(r_begin);
		stash(0);
		__syncthreads();
		int buf = 0;
		for (long long r0 = r_begin; r0 < r_end; r0 += TN_BR) {
			const bool more = r0 + TN_BR < r_end;
			if (more) fetch(r0 + TN_BR);
#pragma unroll
			for (int kk = 0; kk < TN_BR; kk++) {
				float a[TM], b[TN];
#pragma unroll
				for (int i = 0; i < TM; i += 4) {
					if (TM % 4 == 0) {
						const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
						a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
					}
				}
				if (GROUPED) {
#pragma unroll
					for (int j = 0; j < TN; j += 4) {
						const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][(j / 4) * GROUP_STRIDE + tx * 4]);
						b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
					}
				} else {
#pragma unroll
					for (int j = 0; j < TN; j += 2) {
						const float2 t = *reinterpret_cast<const float2*>(&Bs[buf][kk][tx * TN + j]);
						b[j] = t.x; b[j + 1] = t.y;
					}
				}
#pragma unroll
				for (int i = 0; i < TM; i++)
#pragma unroll
					for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
			}
			if (more) stash(buf ^ 1);
			__syncthreads();
			buf ^= 1;
		}
	}
	float* out = partial + (size_t)blockIdx.y * M * N;
#pragma unroll
	for (int i = 0; i < TM; i++) {
		const int gm = m0 + ty * TM + i;
		if (gm >= M) continue;
#pragma unroll
		for (int j = 0; j < TN; j++) {
			const int gn = n0 + column(j);
			if (gn < N) out[(size_t)gm * N + gn] = acc[i][j];
		}
	}
}

// C[i] = sum over splits of partial[s][i], s ascending: the same order on every run and every rank
__global__ void __launch_bounds__(256) matmul_tn_reduce_kernel(const float* __restrict__ partial, float* __restrict__ c, int mn, int splits) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= mn) return;
	float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
	int s = 0;
	for (; s + 4 <= splits; s += 4) {
		s0 += partial[(size_t)(s + 0) * mn + i];
		s1 += partial[(size_t)(s + 1) * mn + i];
		s2 += partial[(size_t)(s + 2) * mn + i];
		s3 += partial[(size_t)(s + 3) * mn + i];
	}
	for (; s < splits; s++) s0 += partial[(size_t)s * mn + i];
	c[i] = (s0 + s1) + (s2 + s3);
}

#ifndef TF_HOST_SIM
// out[M, R] = in[R, M]^T through a 32x33 shared tile; grid.y strides over the rows (used only by the large-output fallback)
__global__ void __launch_bounds__(256) transpose_rm_kernel(const float* __restrict__ in, float* __restrict__ out, long long R, long long M) {
	__shared__ float tile[32][33];
	const long long m0 = (long long)blockIdx.x * 32;
	for (long long r0 = (long long)blockIdx.y * 32; r0 < R; r0 += (long long)gridDim.y * 32) {
		for (int j = threadIdx.y; j < 32; j += 8) {
			const long long rr = r0 + j, mm = m0 + threadIdx.x;
			tile[j][threadIdx.x] = (rr < R && mm < M) ? in[rr * M + mm] : 0.f;
		}
		__syncthreads();
		for (int j = threadIdx.y; j < 32; j += 8) {
			const long long mm = m0 + j, rr = r0 + threadIdx.x;
			if (mm < M && rr < R) out[mm * R + rr] = tile[threadIdx.x][j];
		}
		__syncthreads();
	}
}

template <int BM, int BN, int TM, int TN>
int launch_tn(const float* a, const float* b, float* c, size_t r, size_t m, size_t n) {
	tfcuda::State& s = tfcuda::state();
	const int tiles_m = (int)((m + BM - 1) / BM), tiles_n = (int)((n + BN - 1) / BN);
	const long tiles = (long)tiles_m * tiles_n;
	// split the contraction so that ~4 CTAs per SM are in flight - what the register file holds of the 8x8 tiles (128 registers x 128
	// threads, or 167 x 96 for the 48-row tile), i.e. one wave; every split is a whole number of stages
	constexpr int ctas_per_sm = 4;
	long splits = std::max<long>(1, std::min<long>(((long)s.sm_count * ctas_per_sm + tiles - 1) / tiles, (long)((r + 8 * TN_BR - 1) / (8 * TN_BR))));
	long long rows_per_split = (long long)((r + splits - 1) / splits);
	rows_per_split = (rows_per_split + TN_BR - 1) / TN_BR * TN_BR;
	splits = (long)((r + rows_per_split - 1) / rows_per_split);
	float* partial = static_cast<float*>(tfcuda::scratch((size_t)splits * m * n * sizeof(float)));
	if (!partial) return 1;
	const bool vec = (m % 4 == 0) && (n % 4 == 0) && ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0);
	dim3 grid((unsigned)tiles, (unsigned)splits);
	constexpr int threads = (BM / TM) * (BN / TN);
	{
		tfcuda::ProfileScope prof("lib/matmul_tn", 4.0 * (double)r * (double)(m + n));
		if (vec)
			matmul_tn_kernel<BM, BN, TM, TN, true><<<grid, threads, 0, s.stream>>>(a, b, partial, (long long)r, (int)m, (int)n, rows_per_split, tiles_n);
		else
			matmul_tn_kernel<BM, BN, TM, TN, false><<<grid, threads, 0, s.stream>>>(a, b, partial, (long long)r, (int)m, (int)n, rows_per_split, tiles_n);
		if (tfcuda::check_launch("matmul_tn_kernel")) return 1;
	}
	const int mn = (int)(m * n);
	matmul_tn_reduce_kernel<<<(mn + 255) / 256, 256, 0, s.stream>>>(partial, c, mn, (int)splits);
	if (tfcuda::check_launch("matmul_tn_reduce_kernel")) return 1;
	return 0;
}

#endif  // TF_HOST_SIM

}  // namespace

#ifndef TF_HOST_SIM
extern "C" int tfcuda_matmul_tn(uint64_t a, uint64_t b, uint64_t c, size_t r, size_t m, size_t n) {
	tfcuda::State& s = tfcuda::state();
	if (!s.initialized) { tfcuda::set_error("tfcuda_matmul_tn: not initialised"); return 1; }
	if (m == 0 || n == 0) return 0;
	if (r == 0) return tfcuda_memset32(c, 0, m * n);
	const float* pa = reinterpret_cast<const float*>(a);
	if (m > 0x7fff || n > 0x7fff || m * n > 0x3fffffff) {
		// A weight matrix this large (vocabulary / embedding sized layers) is not the split-K case the kernel below is built for - its
		// partial products alone would be splits x M x N floats.  Materialise A^T once and hand the product to the dense matmul
		// (fp32-accurate 3xTF32 mode, or FFMA where TMA cannot describe the operands): same contract, any extent.
		if (r > 0x7fffffffull || m > 0x7fffffffull) { tfcuda::set_error("tfcuda_matmul_tn: extent exceeds 2^31-1"); return 1; }
		float* at = nullptr;
		TFCUDA_CHECK(cudaMallocAsync(reinterpret_cast<void**>(&at), r * m * sizeof(float), s.stream));
		dim3 grid((unsigned)((m + 31) / 32), (unsigned)std::min<size_t>((r + 31) / 32, 65535));
		{
			tfcuda::ProfileScope prof("lib/matmul_tn_transpose", 8.0 * (double)r * (double)m);
			transpose_rm_kernel<<<grid, dim3(32, 8), 0, s.stream>>>(pa, at, (long long)r, (long long)m);
			if (tfcuda::check_launch("transpose_rm_kernel")) { cudaFreeAsync(at, s.stream); return 1; }
		}
		int rc = tfcuda_matmul(reinterpret_cast<uint64_t>(at), b, c, 1, m, n, r, 1);
		cudaFreeAsync(at, tfcuda::state().stream);
		return rc;
	}
	const float* pb = reinterpret_cast<const float*>(b);
	float* pc = reinterpret_cast<float*>(c);
	// narrow outputs (e.g. the 12 output channels of NCA's second layer) take a tall tile so that lanes are not wasted on padding
	if (n <= 16) return launch_tn<128, 16, 8, 2>(pa, pb, pc, r, m, n);
	if (n <= 32) return launch_tn<128, 32, 8, 4>(pa, pb, pc, r, m, n);
	// a short M (the 48 input channels of NCA's first layer) takes a 48-row tile: the 64-row tile would spend a quarter of its FFMAs on padding
	if (m <= 48) return launch_tn<48, 128, 8, 8>(pa, pb, pc, r, m, n);
	return launch_tn<64, 128, 8, 8>(pa, pb, pc, r, m, n);
}
#endif  // TF_HOST_SIM
