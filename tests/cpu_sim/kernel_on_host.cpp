// TEST INFRASTRUCTURE ONLY.  Compiles hand-written CUDA kernels of tensorfrost_b200/csrc for the HOST through cuda_host_shim.h and runs
// them with one host thread per CUDA thread of a block (blocks one after the other, __syncthreads = std::barrier), against a float64
// reference.  Used for kernels written when no GPU was available (matmul_rows.cu, never run on hardware in round 1); matmul_tn.cu, which
// IS validated on hardware, runs through the same harness as its control.  Checks the kernels' logic - staging, synchronisation
// structure, index arithmetic, tails - not their speed and not the hardware.
//   g++ -std=c++20 -O2 -pthread -I tests/cpu_sim -I tensorfrost_b200/csrc tests/cpu_sim/kernel_on_host.cpp -o kernel_on_host && ./kernel_on_host
#include "cuda_host_shim.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <random>

using std::max;
using std::min;

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
template <typename T> static inline T __ldg(const T* p) { return *p; }
#define __align__(n) __attribute__((aligned(n)))

// warp shuffle stand-in for blocks of ONE warp (launch(..., 32, ...)): every lane publishes its value, the block barrier, every lane
// reads its partner's - all lanes of the warp must call it, as on the device with a full mask
static unsigned sim_shfl_slot[32];
static inline unsigned __shfl_xor_sync(unsigned, unsigned v, int lane_mask) {
	sim_shfl_slot[threadIdx.x] = v;
	__syncthreads();
	const unsigned out = sim_shfl_slot[threadIdx.x ^ (unsigned)lane_mask];
	__syncthreads();
	return out;
}
// the octet-transposed store of the tcgen05 matmul's epilogue (plain CUDA: the part of that kernel that CAN run here)
#include "epilogue_store.cuh"

// matmul_tn.cu declares fixed-size __shared__ arrays inside its kernel: statics shared by the block's threads (the shim's default)
#include "matmul_tn.cu"

// nbody.cu: fixed-size __shared__ tiles too; its PTX wrappers (rsqrt.approx, the packed f32x2 arithmetic) have host stand-ins in the file
struct alignas(16) ulonglong2 { unsigned long long x, y; };
#include "nbody.cu"

// matmul_rows.cu uses dynamic shared memory: `extern __shared__ float smem[]` becomes a reference to this buffer
#undef __shared__
#define __shared__
namespace { alignas(16) float smem[64 * 1024]; }  // the kernel sits in an anonymous namespace: its `extern` declaration resolves here
#include "matmul_rows.cu"

// one block after the other; the threads of a block concurrently, around a barrier
static void launch(unsigned grid_x, unsigned grid_y, unsigned threads, const std::function<void()>& kernel) {
	for (unsigned by = 0; by < grid_y; by++)
		for (unsigned bx = 0; bx < grid_x; bx++) {
			std::barrier<> bar(threads);
			sim_block_barrier = &bar;
			std::vector<std::thread> pool;
			for (unsigned t = 0; t < threads; t++)
				pool.emplace_back([&, t]() {
					gridDim = sim_dim3{grid_x, grid_y, 1};
					blockDim = sim_dim3{threads, 1, 1};
					blockIdx = sim_dim3{bx, by, 0};
					threadIdx = sim_dim3{t, 0, 0};
					kernel();
					bar.arrive_and_drop();
				});
			for (auto& th : pool) th.join();
			sim_block_barrier = nullptr;
		}
}

static std::vector<float> random_matrix(size_t n, unsigned seed) {
	std::mt19937 rng(seed);
	std::normal_distribution<float> dist(0.0f, 1.0f);
	std::vector<float> v(n);
	for (auto& x : v) x = dist(rng);
	return v;
}

static double max_rel_err(const std::vector<float>& got, const std::vector<double>& want) {
	double scale = 1e-30, err = 0;
	for (double w : want) scale = std::max(scale, std::fabs(w));
	for (size_t i = 0; i < want.size(); i++) {
		if (!(got[i] == got[i])) return 1e30;  // NaN: an element was never written
		err = std::max(err, std::fabs((double)got[i] - want[i]));
	}
	return err / scale;
}

template <int TXN, int CH, int TM>
static bool check_rows(size_t r, size_t k, size_t n, unsigned grid) {
	constexpr int BR = (MR_THREADS / TXN) * TM;
	auto a = random_matrix(r * k, (unsigned)(r + k)), b = random_matrix(k * n, (unsigned)(k + n));
	std::vector<float> c(r * n, std::nanf(""));
	std::vector<double> want(r * n, 0.0);
	for (size_t i = 0; i < r; i++)
		for (size_t kk = 0; kk < k; kk++)
			for (size_t j = 0; j < n; j++) want[i * n + j] += (double)a[i * k + kk] * (double)b[kk * n + j];
	const long long tiles = (long long)((r + BR - 1) / BR);
	launch(std::min<unsigned>(grid, (unsigned)tiles), 1, MR_THREADS, [&]() { matmul_rows_kernel<TXN, CH, TM>(a.data(), b.data(), c.data(), (long long)r, (int)k, (int)n, tiles); });
	double e = max_rel_err(c, want);
	std::printf("matmul_rows<%d,%d,%d> R=%zu K=%zu N=%zu grid=%u: max rel err %.2e %s\n", TXN, CH, TM, r, k, n, grid, e, e <= 2e-6 ? "ok" : "FAIL");
	return e <= 2e-6;
}

template <int BM, int BN, int TM, int TN>
static bool check_tn(size_t r, size_t m, size_t n, long splits) {
	auto a = random_matrix(r * m, (unsigned)(r + m)), b = random_matrix(r * n, (unsigned)(r + n));
	std::vector<double> want(m * n, 0.0);
	for (size_t i = 0; i < r; i++)
		for (size_t x = 0; x < m; x++)
			for (size_t y = 0; y < n; y++) want[x * n + y] += (double)a[i * m + x] * (double)b[i * n + y];
	long long rows_per_split = (long long)((r + splits - 1) / splits);
	rows_per_split = (rows_per_split + TN_BR - 1) / TN_BR * TN_BR;
	splits = (long)((r + rows_per_split - 1) / rows_per_split);
	const int tiles_m = (int)((m + BM - 1) / BM), tiles_n = (int)((n + BN - 1) / BN);
	std::vector<float> partial((size_t)splits * m * n, std::nanf("")), c(m * n, std::nanf(""));
	const bool vec = (m % 4 == 0) && (n % 4 == 0);
	constexpr int threads = (BM / TM) * (BN / TN);
	launch((unsigned)(tiles_m * tiles_n), (unsigned)splits, threads, [&]() {
		if (vec) matmul_tn_kernel<BM, BN, TM, TN, true>(a.data(), b.data(), partial.data(), (long long)r, (int)m, (int)n, rows_per_split, tiles_n);
		else matmul_tn_kernel<BM, BN, TM, TN, false>(a.data(), b.data(), partial.data(), (long long)r, (int)m, (int)n, rows_per_split, tiles_n);
	});
	const int mn = (int)(m * n);
	launch((unsigned)((mn + 255) / 256), 1, 256, [&]() { matmul_tn_reduce_kernel(partial.data(), c.data(), mn, (int)splits); });
	double e = max_rel_err(c, want);
	std::printf("matmul_tn<%d,%d,%d,%d> R=%zu M=%zu N=%zu splits=%ld: max rel err %.2e %s\n", BM, BN, TM, TN, r, m, n, splits, e, e <= 2e-6 ? "ok" : "FAIL");
	return e <= 2e-6;
}

// store_block_32x32 against the direct store of the layout it receives (lane = row, r[j] = column j): every element inside m x n written
// with its own value, nothing outside touched (the buffer is NaN-filled with a margin)
static bool check_epilogue_store(int m, int n, int row0, int col0) {
	const int ldc = n, pad_rows = 40;
	std::vector<float> c((size_t)(m + pad_rows) * ldc, std::nanf(""));
	auto value = [&](int row, int col) { return (float)(row * 1000 + col) + 0.25f; };
	launch(1, 1, 32, [&]() {
		const int lane = (int)threadIdx.x;
		uint32_t r[32];
		for (int j = 0; j < 32; j++) r[j] = __float_as_uint(value(row0 + lane, col0 + j));
		store_block_32x32(r, c.data(), (size_t)ldc, row0, col0, m, n, lane);
	});
	bool good = true;
	for (int row = 0; row < m + pad_rows && good; row++)
		for (int col = 0; col < ldc; col++) {
			const bool inside = row >= row0 && row < row0 + 32 && row < m && col >= col0 && col < col0 + 32 && col < n;
			const float got = c[(size_t)row * ldc + col];
			if (inside ? got != value(row, col) : got == got) { good = false; break; }
		}
	std::printf("epilogue store m=%d n=%d block at (%d, %d): %s\n", m, n, row0, col0, good ? "ok" : "FAIL");
	return good;
}

// one gravity step: both kernels (scalar loop; packed f32x2 loop with its padded tail tile) against a float64 evaluation of the same formula
static bool check_nbody(int n, bool packed) {
	auto x = random_matrix((size_t)n * 3, (unsigned)n);
	for (auto& e : x) e *= 5.0f;
	auto v = random_matrix((size_t)n * 3, (unsigned)n + 1);
	for (auto& e : v) e *= 0.1f;
	std::vector<float> xn((size_t)n * 3, std::nanf("")), vn((size_t)n * 3, std::nanf(""));
	const float dt = 0.001f, eps = 1e-4f;
	std::vector<double> want_v((size_t)n * 3), want_x((size_t)n * 3);
	for (int i = 0; i < n; i++) {
		double f[3] = {0, 0, 0};
		for (int j = 0; j < n; j++) {
			double d[3] = {(double)x[3 * i] - x[3 * j], (double)x[3 * i + 1] - x[3 * j + 1], (double)x[3 * i + 2] - x[3 * j + 2]};
			double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2] + eps;
			double w = 1.0 / (d2 * std::sqrt(d2));
			for (int c = 0; c < 3; c++) f[c] -= d[c] * w;
		}
		for (int c = 0; c < 3; c++) {
			want_v[3 * i + c] = v[3 * i + c] + f[c] * dt;
			want_x[3 * i + c] = x[3 * i + c] + want_v[3 * i + c] * dt;
		}
	}
	if (packed)
		launch((unsigned)((n + NX_THREADS * 2 - 1) / (NX_THREADS * 2)), 1, NX_THREADS, [&]() { nbody_kernel_x2(x.data(), v.data(), xn.data(), vn.data(), n, dt, eps); });
	else
		launch((unsigned)((n + NB_THREADS * NB_PER_THREAD - 1) / (NB_THREADS * NB_PER_THREAD)), 1, NB_THREADS, [&]() { nbody_kernel(x.data(), v.data(), xn.data(), vn.data(), n, dt, eps); });
	double ev = max_rel_err(vn, want_v), ex = max_rel_err(xn, want_x);
	bool good = ev <= 1e-5 && ex <= 1e-6;
	std::printf("nbody %s n=%d: v max rel err %.2e, x %.2e %s\n", packed ? "packed f32x2" : "scalar", n, ev, ex, good ? "ok" : "FAIL");
	return good;
}

int main() {
	bool ok = true;
	// the n-body step: scalar kernel (validated on hardware through tests/test_library_gpu.py) and the packed kernel at the two sizes of
	// tests/test_zy_late_gpu.py (5000 = nine full 512-body tiles + a padded tail)
	ok &= check_nbody(1500, false);
	ok &= check_nbody(4096, true);
	ok &= check_nbody(5000, true);
	// the tcgen05 matmul's epilogue store (octet transpose by shuffles): full block, ragged rows, ragged columns (float4 and scalar tails),
	// a block entirely outside, N smaller than a block
	ok &= check_epilogue_store(64, 128, 32, 96);
	ok &= check_epilogue_store(50, 128, 32, 0);
	ok &= check_epilogue_store(64, 44, 0, 32);
	ok &= check_epilogue_store(70, 12, 64, 0);
	ok &= check_epilogue_store(20, 64, 32, 32);
	// control: the hardware-validated weight-gradient kernel through the same harness
	ok &= check_tn<64, 128, 8, 8>(1000, 48, 128, 5);
	ok &= check_tn<64, 128, 8, 8>(500, 100, 200, 3);   // several tiles in both directions, ragged in both (grouped column layout of the 8-wide tile)
	ok &= check_tn<48, 128, 8, 8>(1000, 48, 128, 5);   // the 48-row tile NCA's first layer takes (96 threads)
	ok &= check_tn<48, 128, 8, 8>(333, 37, 130, 4);    // M and N not multiples of 4: scalar loads, two column tiles, ragged tails
	ok &= check_tn<128, 16, 8, 2>(777, 128, 12, 3);
	ok &= check_tn<128, 32, 8, 4>(300, 20, 24, 2);
	// the skinny matmul written without a GPU: every template configuration the dispatcher uses, NCA's four shapes, ragged tails
	ok &= check_rows<16, 2, 4>(200, 48, 128, 2);    // fc1 forward: K = 48 (two chunks: 32 + 16), BN 128
	ok &= check_rows<4, 1, 4>(600, 128, 12, 2);     // fc2 forward: K = 128 (four chunks), N = 12 in a 16-wide tile
	ok &= check_rows<16, 2, 4>(130, 12, 128, 3);    // dX2: K = 12 (one short chunk)
	ok &= check_rows<16, 1, 4>(100, 128, 48, 1);    // dX1: N = 48 in a 64-wide tile
	ok &= check_rows<8, 1, 4>(257, 36, 30, 2);      // BN 32, N not a multiple of 4 (scalar stores), ragged last tile
	ok &= check_rows<4, 1, 4>(300, 7, 5, 1);        // K not a multiple of 4: scalar A loads, zero-padded k
	ok &= check_rows<16, 2, 4>(1, 4, 128, 4);       // one row, more CTAs than tiles
	std::printf(ok ? "ALL OK\n" : "SOME FAILED\n");
	return ok ? 0 : 1;
}
