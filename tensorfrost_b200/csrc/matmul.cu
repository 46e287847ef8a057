// C[b] = A[b] (M x K) @ B[b] (K x N), row-major fp32.
//
// Replaces ComputeMatMul (TensorFrost/Compiler/Implementations.cpp:560-646): one thread per C element with a
// serial k loop and an uncoalesced walk down B's column ("very much not optimized", README.md:820).
//   mode 2 (this file): fp32 FFMA kernel — 128x128 CTA tile, 8x8 register tile per thread, K staged through
//       double-buffered shared memory with 128-bit loads.  Same products and the same fp32 accumulator as the
//       oracle, summed in k order per element (bit-compatible up to FMA contraction).
//   mode 0 / 1: tcgen05 kind::tf32 with TMEM accumulators fed by TMA (matmul_tcgen05.cu).
#include "tfcuda_internal.h"

int tfcuda_matmul_tcgen05(const float* a, const float* b, float* c, size_t batch, size_t m, size_t n, size_t k, int mode);
bool tfcuda_matmul_tcgen05_supported(const float* a, const float* b, const float* c, size_t m, size_t n, size_t k);

namespace {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
constexpr int MM_THREADS = (BM / TM) * (BN / TN);  // 256

__global__ void __launch_bounds__(MM_THREADS) matmul_ffma_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int M, int N, int K) {
	__shared__ float As[2][BK][BM + 4];  // A tile stored transposed: As[k][m]
	__shared__ float Bs[2][BK][BN];
	const size_t batch = blockIdx.z;
	A += batch * (size_t)M * K;
	B += batch * (size_t)K * N;
	C += batch * (size_t)M * N;
	const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
	const int tid = threadIdx.x;
	const int tx = tid % (BN / TN), ty = tid / (BN / TN);

	float acc[TM][TN];
#pragma unroll
	for (int i = 0; i < TM; i++)
#pragma unroll
		for (int j = 0; j < TN; j++) acc[i][j] = 0.0f;

	// loaders: A tile 128x16 -> 2048 floats, 8 per thread; B tile 16x128 -> 8 per thread
	auto load_tiles = [&](int buf, int k0) {
#pragma unroll
		for (int r = 0; r < (BM * BK) / MM_THREADS; r++) {
			int e = tid + r * MM_THREADS;
			int mm = e / BK, kk = e % BK;
			int gm = m0 + mm, gk = k0 + kk;
			As[buf][kk][mm] = (gm < M && gk < K) ? A[(size_t)gm * K + gk] : 0.0f;
		}
#pragma unroll
		for (int r = 0; r < (BK * BN) / MM_THREADS; r++) {
			int e = tid + r * MM_THREADS;
			int kk = e / BN, nn = e % BN;
			int gk = k0 + kk, gn = n0 + nn;
			Bs[buf][kk][nn] = (gk < K && gn < N) ? B[(size_t)gk * N + gn] : 0.0f;
		}
	};

	load_tiles(0, 0);
	__syncthreads();
	int buf = 0;
	for (int k0 = 0; k0 < K; k0 += BK) {
		if (k0 + BK < K) load_tiles(buf ^ 1, k0 + BK);
#pragma unroll
		for (int kk = 0; kk < BK; kk++) {
			float a[TM], b[TN];
#pragma unroll
			for (int i = 0; i < TM; i++) a[i] = As[buf][kk][ty * TM + i];
#pragma unroll
			for (int j = 0; j < TN; j++) b[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
			for (int i = 0; i < TM; i++)
#pragma unroll
				for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
		}
		__syncthreads();
		buf ^= 1;
	}
#pragma unroll
	for (int i = 0; i < TM; i++) {
		int gm = m0 + ty * TM + i;
		if (gm >= M) continue;
#pragma unroll
		for (int j = 0; j < TN; j++) {
			int gn = n0 + tx * TN + j;
			if (gn < N) C[(size_t)gm * N + gn] = acc[i][j];
		}
	}
}

}  // namespace

extern "C" int tfcuda_matmul(uint64_t a, uint64_t b, uint64_t c, size_t batch, size_t m, size_t n, size_t k, int mode) {
	tfcuda::State& s = tfcuda::state();
	if (!s.initialized) { tfcuda::set_error("tfcuda_matmul: not initialised"); return 1; }
	if (batch == 0 || m == 0 || n == 0) return 0;
	if (k == 0) return tfcuda_memset32(c, 0, batch * m * n);
	if (m > 0x7fffffff || n > 0x7fffffff || k > 0x7fffffff || batch > 65535) { tfcuda::set_error("tfcuda_matmul: extent out of range"); return 1; }
	const float* pa = reinterpret_cast<const float*>(a);
	const float* pb = reinterpret_cast<const float*>(b);
	float* pc = reinterpret_cast<float*>(c);
	if (mode < 0 || mode > 2) { tfcuda::set_error("tfcuda_matmul: unknown mode"); return 1; }
	// modes 0/1 need TMA-describable operands (16-byte aligned bases and row pitches, every batch slice); other shapes take the FFMA kernel
	if (mode != 2 && tfcuda_matmul_tcgen05_supported(pa, pb, pc, m, n, k) && (batch == 1 || ((m * k) % 4 == 0 && (k * n) % 4 == 0 && (m * n) % 4 == 0)))
		return tfcuda_matmul_tcgen05(pa, pb, pc, batch, m, n, k, mode);
	tfcuda::ProfileScope prof("lib/matmul_ffma");
	dim3 grid((unsigned)((n + BN - 1) / BN), (unsigned)((m + BM - 1) / BM), (unsigned)batch);
	matmul_ffma_kernel<<<grid, MM_THREADS, 0, s.stream>>>(pa, pb, pc, (int)m, (int)n, (int)k);
	return tfcuda::check_launch("tfcuda_matmul(ffma)");
}
