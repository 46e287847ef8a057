// tcgen05 / TMEM / TMA matmul: C = A (M x K, row-major fp32) @ B (K x N, row-major fp32), modes 0 and 1 of tfcuda_matmul.
//
// Replaces the reference's generic lowering of `matmul` (TensorFrost/Compiler/Implementations.cpp:560-646: one thread per C
// element, serial k loop).  The contraction is the one genuinely dense GEMM of the path, so it runs on the 5th-generation
// tensor cores:
//   * operands stay fp32 in HBM and are fed as kind::tf32 (the tensor core reads the top 19 bits of each word);
//   * both operand tiles are K-major in shared memory: A [128 x 32] as it lies in memory, B from a transposed copy Bt [N x K]
//     made by a tiled pre-pass (measured on B200: kind::tf32 with an MN-major B descriptor silently yields a zero accumulator,
//     tools/tcgen05_probe.cu, so the N-major B of a row-major matmul cannot be fed directly).  TMA (cp.async.bulk.tensor.2d,
//     SWIZZLE_128B) drops the tiles into a shared-memory ring, one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=8
//     per instruction) with the accumulator in TMEM, and 4 epilogue warps drain TMEM with tcgen05.ld straight to global memory;
//   * persistent CTAs (one per SM) walk a grouped tile order so concurrently running tiles share A row-blocks / B column-blocks
//     in L2; TMEM holds two accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.
// mode 0: one TF32 product per k-step (1e-3 class: 10-bit mantissa operands, fp32 accumulate).
// mode 1: error-compensated 3xTF32: x = hi + lo with hi = x & 0xffffe000 (exact in tf32) and lo = x - hi; the kernel accumulates
//         Ahi*Bhi + Ahi*Blo + Alo*Bhi (the dropped lo*lo term is 2^-22 relative), which restores fp32-level accuracy (~1e-6).
//         The hi/lo planes come from the pre-pass (split for A, transpose+split for B), independent of how the hardware rounds.
// Shapes TMA cannot describe (row pitch not a multiple of 16 bytes) are served by the FFMA kernel in matmul.cu.
#include <cuda.h>

#include "tfcuda_internal.h"
#include "epilogue_store.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;      // 32 fp32 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 8;        // kind::tf32
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KB
constexpr int NUM_THREADS = 192;                     // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
constexpr int GROUP_M = 16;

// ---- PTX wrappers ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "WAIT_LOOP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra WAIT_DONE;\n\t"
	    "bra WAIT_LOOP;\n\t"
	    "WAIT_DONE:\n\t"
	    "}" ::"r"(bar), "r"(parity)
	    : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
	asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map), "r"(bar),
	             "r"(c0), "r"(c1)
	             : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
	    "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
	    : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
	asm volatile(
	    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
	    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
	    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
	    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
	      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
	      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
	      "=r"(r[31])
	    : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor (SM100 UMMA): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | version 1 <<46 | SWIZZLE_128B (2) <<61
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
	d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
	d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
	d |= (uint64_t)1 << 46;
	d |= (uint64_t)2 << 61;
	return d;
}

// instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), A and B K-major (bits 15, 16 = 0), N>>3 <<17, M>>4 <<24
__host__ __device__ constexpr uint32_t make_idesc(int n) {
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

struct GemmArgs {
	unsigned* dbg;  // TFCUDA_GEMM_DEBUG: CTA 0 dumps its TMEM base, the head of the first A/B stage and of the first accumulator row
	float* c;
	int m, n, k;
	int tiles_m, tiles_n;
};

template <int BLOCK_N, bool SPLIT3>
struct Config {
	static constexpr int kPlanes = SPLIT3 ? 2 : 1;
	static constexpr int kBBytes = BLOCK_N * BLOCK_K * 4;  // [BLOCK_N rows of Bt][128 B]
	static constexpr int kStageBytes = kPlanes * (A_TILE_BYTES + kBBytes);
	static constexpr int kStages = (200 * 1024 / kStageBytes) > 8 ? 8 : (200 * 1024 / kStageBytes);
	static constexpr int kTmemCols = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
	static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*alignment slack*/ + 256 /*barriers*/;
	static_assert(kStages >= 2, "pipeline too shallow");
};

__device__ __forceinline__ void tile_coords(int tile, int tiles_m, int tiles_n, int& m_blk, int& n_blk) {
	// grouped order: GROUP_M row-blocks share their B column-blocks while they are hot in L2
	const int per_group = GROUP_M * tiles_n;
	const int group = tile / per_group;
	const int first_m = group * GROUP_M;
	const int rows = min(GROUP_M, tiles_m - first_m);
	const int in_group = tile - group * per_group;
	m_blk = first_m + in_group % rows;
	n_blk = in_group / rows;
}

template <int BLOCK_N, bool SPLIT3>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_lo, const __grid_constant__ GemmArgs g) {
	using Cfg = Config<BLOCK_N, SPLIT3>;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t bar_base = smem_base + Cfg::kStages * Cfg::kStageBytes;
	// barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base address word
	auto full_bar = [&](int s) { return bar_base + 8u * s; };
	auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::kStages + s); };
	auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + s); };
	auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::kStages + 2 + s); };
	const uint32_t tmem_slot = bar_base + 8u * (2 * Cfg::kStages + 4);

	const int warp = threadIdx.x >> 5;
	const int lane = threadIdx.x & 31;
	const int num_tiles = g.tiles_m * g.tiles_n;
	const int num_k_blocks = (g.k + BLOCK_K - 1) / BLOCK_K;

	if (warp == 0 && lane == 0) {
		for (int s = 0; s < Cfg::kStages; s++) {
			mbar_init(full_bar(s), 1);
			mbar_init(empty_bar(s), 1);
		}
		for (int s = 0; s < 2; s++) {
			mbar_init(tmem_full_bar(s), 1);
			mbar_init(tmem_empty_bar(s), 4);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)Cfg::kTmemCols) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	tcgen05_fence_before();
	__syncthreads();
	tcgen05_fence_after();
	uint32_t tmem_base;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

	if (warp == 0) {
		// ===== TMA producer =====
		if (lane == 0) {
			int stage = 0;
			uint32_t phase = 0;
			for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
				int m_blk, n_blk;
				tile_coords(tile, g.tiles_m, g.tiles_n, m_blk, n_blk);
				for (int kb = 0; kb < num_k_blocks; kb++) {
					mbar_wait(empty_bar(stage), phase ^ 1);
					const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
					const uint32_t sb = sa + Cfg::kPlanes * A_TILE_BYTES;
					mbar_expect_tx(full_bar(stage), Cfg::kStageBytes);
					tma_load_2d(sa, &map_a, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
					if (SPLIT3) tma_load_2d(sa + A_TILE_BYTES, &map_a_lo, full_bar(stage), kb * BLOCK_K, m_blk * BLOCK_M);
					tma_load_2d(sb, &map_b, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
					if (SPLIT3) tma_load_2d(sb + Cfg::kBBytes, &map_b_lo, full_bar(stage), kb * BLOCK_K, n_blk * BLOCK_N);
					if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
				}
			}
		}
	} else if (warp == 1) {
		// ===== MMA issuer (one thread) =====
		if (lane == 0) {
			constexpr uint32_t idesc = make_idesc(BLOCK_N);
			int stage = 0;
			uint32_t phase = 0;
			int acc = 0;
			uint32_t acc_phase = 0;
			for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
				mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
				tcgen05_fence_after();
				const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
				for (int kb = 0; kb < num_k_blocks; kb++) {
					mbar_wait(full_bar(stage), phase);
					tcgen05_fence_after();
					const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
					const uint32_t sb = sa + Cfg::kPlanes * A_TILE_BYTES;
					if (g.dbg && blockIdx.x == 0 && tile == 0 && kb == 0) {
						g.dbg[0] = tmem_base;
						g.dbg[1] = smem_base;
						for (int i = 0; i < 64; i++) {
							uint32_t va, vb;
							asm volatile("ld.shared.b32 %0, [%1];" : "=r"(va) : "r"(sa + 4 * i));
							asm volatile("ld.shared.b32 %0, [%1];" : "=r"(vb) : "r"(sb + 4 * i));
							g.dbg[16 + i] = va;
							g.dbg[80 + i] = vb;
						}
					}
#pragma unroll
					for (int ks = 0; ks < BLOCK_K / UMMA_K; ks++) {
						// K-major operands: 8-row groups 1024 B apart (SBO); one k-step = 32 bytes along the swizzled 128-byte row
						const uint64_t da = make_desc(sa + ks * UMMA_K * 4, 16, 1024);
						const uint64_t db = make_desc(sb + ks * UMMA_K * 4, 16, 1024);
						umma_tf32(tmem_d, da, db, idesc, (kb | ks) != 0);
						if (SPLIT3) {
							const uint64_t da_lo = make_desc(sa + A_TILE_BYTES + ks * UMMA_K * 4, 16, 1024);
							const uint64_t db_lo = make_desc(sb + Cfg::kBBytes + ks * UMMA_K * 4, 16, 1024);
							umma_tf32(tmem_d, da, db_lo, idesc, 1);
							umma_tf32(tmem_d, da_lo, db, idesc, 1);
						}
					}
					umma_commit(empty_bar(stage));  // frees the smem slot when the MMAs above have read it
					if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
				}
				umma_commit(tmem_full_bar(acc));  // accumulator complete
				if (++acc == 2) { acc = 0; acc_phase ^= 1; }
			}
		}
	} else {
		// ===== epilogue: TMEM -> registers -> global =====
		const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter + 32) belong to this warp
		int acc = 0;
		uint32_t acc_phase = 0;
		for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
			int m_blk, n_blk;
			tile_coords(tile, g.tiles_m, g.tiles_n, m_blk, n_blk);
			mbar_wait(tmem_full_bar(acc), acc_phase);
			tcgen05_fence_after();
#pragma unroll 1
			for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
				uint32_t r[32];
				tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + c0, r);
				tmem_ld_wait();
				if (g.dbg && blockIdx.x == 0 && tile == 0 && c0 == 0 && lane == 0) {
					for (int i = 0; i < 8; i++) g.dbg[160 + quarter * 8 + i] = r[i];
				}
				// octet-transposed store: every store instruction of the warp covers 4 rows x 128 contiguous bytes (epilogue_store.cuh)
				store_block_32x32(r, g.c, (size_t)g.n, m_blk * BLOCK_M + quarter * 32, n_blk * BLOCK_N + c0, g.m, g.n, lane);
			}
			tcgen05_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
			if (++acc == 2) { acc = 0; acc_phase ^= 1; }
		}
	}

	tcgen05_fence_before();
	__syncthreads();
	if (warp == 1) {
		tcgen05_fence_after();
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols) : "memory");
	}
}

// hi = x with the 13 low mantissa bits cleared (exactly representable in tf32), lo = x - hi (exact in fp32).  For a non-finite x
// (Inf - Inf would make lo NaN and poison whole rows where the reference's fp32 loop yields Inf) lo is 0.  The tensor core truncates lo
// to 10 mantissa bits again, so the compensated product is accurate to ~2^-21 relative, not the full 2^-24 of an fp32 FMA chain.
__device__ __forceinline__ float tf32_lo(float v, float h) { return (__float_as_uint(h) & 0x7f800000u) == 0x7f800000u ? 0.f : v - h; }

__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	const size_t n4 = n >> 2;
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += stride) {
		float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
		float4 h, l;
		h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = tf32_lo(v.x, h.x);
		h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = tf32_lo(v.y, h.y);
		h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = tf32_lo(v.z, h.z);
		h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = tf32_lo(v.w, h.w);
		reinterpret_cast<float4*>(hi)[i] = h;
		reinterpret_cast<float4*>(lo)[i] = l;
	}
	for (size_t i = (n4 << 2) + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
		float v = x[i];
		float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
		hi[i] = h;
		lo[i] = tf32_lo(v, h);
	}
}

// Bt[n][k] = B[k][n] through 32x32 shared-memory tiles (coalesced on both sides); with SPLIT also the hi/lo planes of Bt
template <bool SPLIT>
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ b, float* __restrict__ bt, float* __restrict__ bt_lo, int k, int n) {
	__shared__ float tile[32][33];
	const int tiles_n = (n + 31) / 32, tiles_k = (k + 31) / 32;
	const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
	for (int t = blockIdx.x; t < tiles_n * tiles_k; t += gridDim.x) {
		const int k0 = (t / tiles_n) * 32, n0 = (t % tiles_n) * 32;
#pragma unroll
		for (int r = ty; r < 32; r += 8) tile[r][tx] = (k0 + r < k && n0 + tx < n) ? b[(size_t)(k0 + r) * n + n0 + tx] : 0.0f;
		__syncthreads();
#pragma unroll
		for (int r = ty; r < 32; r += 8) {
			if (n0 + r < n && k0 + tx < k) {
				float v = tile[tx][r];
				if (SPLIT) {
					float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
					bt[(size_t)(n0 + r) * k + k0 + tx] = h;
					bt_lo[(size_t)(n0 + r) * k + k0 + tx] = tf32_lo(v, h);
				} else {
					bt[(size_t)(n0 + r) * k + k0 + tx] = v;
				}
			}
		}
		__syncthreads();
	}
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
	static EncodeTiledFn fn = nullptr;
	if (!fn) {
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
	}
	return fn;
}

// 2-D fp32 tensor [rows][cols] row-major, box [box_rows][32 floats], 128-byte swizzle, out-of-bounds elements read as 0
bool make_map(CUtensorMap* map, const float* base, size_t rows, size_t cols, unsigned box_rows) {
	EncodeTiledFn fn = encode_fn();
	if (!fn) {
		tfcuda::set_error("cuTensorMapEncodeTiled is not available in this driver");
		return false;
	}
	cuuint64_t dims[2] = {cols, rows};
	cuuint64_t strides[1] = {cols * sizeof(float)};
	cuuint32_t box[2] = {32, box_rows};
	cuuint32_t elem[2] = {1, 1};
	CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
	                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		tfcuda::set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " for a [" + std::to_string(rows) + " x " + std::to_string(cols) + "] tensor");
		return false;
	}
	return true;
}

template <int BLOCK_N, bool SPLIT3>
int launch(const float* a, const float* b, const float* a_lo, const float* b_lo, float* c, size_t m, size_t n, size_t k) {
	using Cfg = Config<BLOCK_N, SPLIT3>;
	tfcuda::State& s = tfcuda::state();
	CUtensorMap ma, mb, mal, mbl;
	// b / b_lo are the TRANSPOSED planes Bt [n][k]
	if (!make_map(&ma, a, m, k, BLOCK_M) || !make_map(&mb, b, n, k, BLOCK_N)) return 1;
	if (SPLIT3) {
		if (!make_map(&mal, a_lo, m, k, BLOCK_M) || !make_map(&mbl, b_lo, n, k, BLOCK_N)) return 1;
	} else {
		mal = ma;
		mbl = mb;
	}
	GemmArgs g;
	g.dbg = nullptr;
	static const bool debug = getenv("TFCUDA_GEMM_DEBUG") != nullptr;
	if (debug) {
		TFCUDA_CHECK(cudaMalloc(&g.dbg, 256 * 4));
		TFCUDA_CHECK(cudaMemset(g.dbg, 0xff, 256 * 4));
	}
	g.c = c;
	g.m = (int)m;
	g.n = (int)n;
	g.k = (int)k;
	g.tiles_m = (int)((m + BLOCK_M - 1) / BLOCK_M);
	g.tiles_n = (int)((n + BLOCK_N - 1) / BLOCK_N);
	static bool attr_set = false;
	if (!attr_set) {
		TFCUDA_CHECK(cudaFuncSetAttribute(gemm_tf32_kernel<BLOCK_N, SPLIT3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
		attr_set = true;
	}
	const long tiles = (long)g.tiles_m * g.tiles_n;
	const unsigned grid = (unsigned)std::min<long>(tiles, s.sm_count);
	gemm_tf32_kernel<BLOCK_N, SPLIT3><<<grid, NUM_THREADS, Cfg::kSmemBytes, s.stream>>>(ma, mb, mal, mbl, g);
	if (debug) {
		unsigned h[256];
		cudaError_t e = cudaStreamSynchronize(s.stream);
		cudaMemcpy(h, g.dbg, sizeof(h), cudaMemcpyDeviceToHost);
		cudaFree(g.dbg);
		fprintf(stderr, "[gemm debug] BLOCK_N=%d split=%d sync=%s tmem_base=%08x smem_base=%08x\n", BLOCK_N, (int)SPLIT3, cudaGetErrorName(e), h[0], h[1]);
		fprintf(stderr, "  A smem:");
		for (int i = 0; i < 64; i++) fprintf(stderr, " %g", (double)*reinterpret_cast<float*>(&h[16 + i]));
		fprintf(stderr, "\n  B smem:");
		for (int i = 0; i < 64; i++) fprintf(stderr, " %g", (double)*reinterpret_cast<float*>(&h[80 + i]));
		fprintf(stderr, "\n  acc row0 per quarter:");
		for (int i = 0; i < 32; i++) fprintf(stderr, " %g", (double)*reinterpret_cast<float*>(&h[160 + i]));
		fprintf(stderr, "\n");
	}
	return tfcuda::check_launch(SPLIT3 ? "gemm_tf32_kernel(3xTF32)" : "gemm_tf32_kernel");
}

}  // namespace

bool tfcuda_matmul_tcgen05_supported(const float* a, const float* b, const float* c, size_t m, size_t n, size_t k) {
	// TMA needs 16-byte aligned bases and row pitches (A [m][k] and the transposed copy Bt [n][k]: k % 4); C rows are stored with float4 (n % 4)
	return ((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0) && ((uintptr_t)c % 16 == 0) && (k % 4 == 0) && (n % 4 == 0) && m > 0 && n > 0 && k > 0 &&
	       m < (1u << 31) && n < (1u << 31) && k < (1u << 31);
}

int tfcuda_matmul_tcgen05(const float* a, const float* b, float* c, size_t batch, size_t m, size_t n, size_t k, int mode) {
	tfcuda::State& s = tfcuda::state();
	const size_t na = m * k, nb = k * n;
	const size_t na4 = (na + 3) & ~size_t(3), nb4 = (nb + 3) & ~size_t(3);
	// scratch planes: mode 0: [Bt]; mode 1: [A_hi | A_lo | Bt_hi | Bt_lo]
	float* scratch = static_cast<float*>(tfcuda::scratch((mode == 0 ? nb4 : 2 * na4 + 2 * nb4) * sizeof(float)));
	if (!scratch) return 1;
	const unsigned t_tiles = (unsigned)(((n + 31) / 32) * ((k + 31) / 32));
	const unsigned t_grid = std::min<unsigned>(t_tiles, (unsigned)s.sm_count * 16);
	int rc = 0;
	for (size_t bi = 0; bi < batch && rc == 0; bi++) {
		const float* pa = a + bi * na;
		const float* pb = b + bi * nb;
		float* pc = c + bi * m * n;
		if (mode == 0) {
			float* bt = scratch;
			{
				tfcuda::ProfileScope prof("lib/matmul_transpose_b");
				transpose_kernel<false><<<t_grid, 256, 0, s.stream>>>(pb, bt, nullptr, (int)k, (int)n);
				if ((rc = tfcuda::check_launch("transpose_kernel"))) break;
			}
			tfcuda::ProfileScope prof("lib/matmul_tcgen05_tf32");
			if (n > 128) rc = launch<256, false>(pa, bt, nullptr, nullptr, pc, m, n, k);
			else if (n > 64) rc = launch<128, false>(pa, bt, nullptr, nullptr, pc, m, n, k);
			else if (n > 32) rc = launch<64, false>(pa, bt, nullptr, nullptr, pc, m, n, k);
			else rc = launch<32, false>(pa, bt, nullptr, nullptr, pc, m, n, k);
		} else {
			float* a_hi = scratch;
			float* a_lo = a_hi + na4;
			float* bt_hi = a_lo + na4;
			float* bt_lo = bt_hi + nb4;
			{
				tfcuda::ProfileScope prof("lib/matmul_split_tf32");
				unsigned ga = (unsigned)std::min<size_t>((na / 4 + 255) / 256 + 1, (size_t)s.sm_count * 8);
				split_tf32_kernel<<<ga, 256, 0, s.stream>>>(pa, a_hi, a_lo, na);
				if ((rc = tfcuda::check_launch("split_tf32_kernel"))) break;
				transpose_kernel<true><<<t_grid, 256, 0, s.stream>>>(pb, bt_hi, bt_lo, (int)k, (int)n);
				if ((rc = tfcuda::check_launch("transpose_kernel"))) break;
			}
			tfcuda::ProfileScope prof("lib/matmul_tcgen05_3xtf32");
			if (n > 64) rc = launch<128, true>(a_hi, bt_hi, a_lo, bt_lo, pc, m, n, k);
			else if (n > 32) rc = launch<64, true>(a_hi, bt_hi, a_lo, bt_lo, pc, m, n, k);
			else rc = launch<32, true>(a_hi, bt_hi, a_lo, bt_lo, pc, m, n, k);
		}
	}
	return rc;
}
