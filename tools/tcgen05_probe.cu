// Stand-alone probe of one tcgen05.mma kind::tf32 instruction group (development tool, not part of the product):
// one CTA loads A[128x32] and B via TMA, issues 4 MMAs (K=8 each), reads the accumulator back and compares with the host.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tcgen05_probe tools/tcgen05_probe.cu -lcuda
//   run:   build/tcgen05_probe <variant>      variant bit0: B K-major (B given as [N][K]); bit1: disable-lane operand form;
//                                              bit2: A/B desc version bits 0; bit3: swap LBO/SBO of the MN-major B; bit4: LBO-mode bit
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	asm volatile(
	    "{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra WD;\n\tbra WL;\n\tWD:\n\t}" ::"r"(bar), "r"(parity)
	    : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int version) {
	uint64_t d = 0;
	d |= (uint64_t)((addr >> 4) & 0x3fff);
	d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
	d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
	d |= (uint64_t)version << 46;
	d |= (uint64_t)2 << 61;
	return d;
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* c, int variant,
                                                unsigned* dbg) {
	extern __shared__ uint8_t raw[];
	const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
	const uint32_t sa = base;                   // 128 x 128 B = 16 KB
	const uint32_t sb = base + 16384;           // up to 64 x 128 B = 8 KB
	const uint32_t bar_full = base + 32768, bar_mma = bar_full + 8, slot = bar_full + 16;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const bool b_kmajor = variant & 1;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_mma));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(slot) : "memory");
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	uint32_t tmem;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));
	if (threadIdx.x == 0) {
		const uint32_t bytes = 16384 + 8192;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_full), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sa), "l"(&map_a),
		             "r"(bar_full), "r"(0), "r"(0)
		             : "memory");
		if (b_kmajor) {
			// B^T as [N=64][K=32]: one box of 64 rows x 128 B
			asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sb), "l"(&map_b),
			             "r"(bar_full), "r"(0), "r"(0)
			             : "memory");
		} else {
			// B as [K=32][N=64]: two boxes of 32 k-rows x 32 n
			for (int j = 0; j < 2; j++)
				asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sb + j * 4096),
				             "l"(&map_b), "r"(bar_full), "r"(j * 32), "r"(0)
				             : "memory");
		}
		mbar_wait(bar_full, 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		const int version = (variant & 4) ? 0 : 1;
		const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((b_kmajor ? 0u : 1u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
		dbg[0] = tmem;
		dbg[1] = idesc;
		for (int ks = 0; ks < 4; ks++) {
			uint64_t da = make_desc(sa + ks * 32, 16, 1024, version);
			uint64_t db = b_kmajor ? make_desc(sb + ks * 32, 16, 1024, version) : ((variant & 8) ? make_desc(sb + ks * 1024, 1024, 4096, version) : make_desc(sb + ks * 1024, 4096, 1024, version));
			if (variant & 16) db |= (uint64_t)1 << 52;  // LBO mode bit
			uint32_t acc = ks != 0;
			if (variant & 2) {
				uint32_t z = 0;
				asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(tmem),
				             "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(z)
				             : "memory");
			} else {
				asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da),
				             "l"(db), "r"(idesc), "r"(acc)
				             : "memory");
			}
		}
		asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_mma) : "memory");
	}
	mbar_wait(bar_mma, 0);
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	for (int c0 = 0; c0 < N; c0 += 32) {
		uint32_t r[32];
		asm volatile(
		    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
		    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
		    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
		    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
		      "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
		      "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
		      "=r"(r[31])
		    : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		for (int j = 0; j < 32; j++) c[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
	int variant = argc > 1 ? atoi(argv[1]) : 0;
	const bool b_kmajor = variant & 1;
	cudaFree(0);
	void* fp = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
	EncodeFn enc = (EncodeFn)fp;
	std::vector<float> a(M * K), b(K * N), bt(N * K), cref(M * N, 0.f), c(M * N, -1.f);
	srand(1);
	for (auto& v : a) v = (float)(rand() % 16) / 4.f;   // exactly representable in tf32
	for (auto& v : b) v = (float)(rand() % 16) / 8.f;
	for (int k = 0; k < K; k++)
		for (int n = 0; n < N; n++) bt[n * K + k] = b[k * N + n];
	for (int i = 0; i < M; i++)
		for (int n = 0; n < N; n++) {
			float s = 0;
			for (int k = 0; k < K; k++) s += a[i * K + k] * b[k * N + n];
			cref[i * N + n] = s;
		}
	float *da, *db, *dc;
	unsigned* dbg;
	cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dc, c.size() * 4); cudaMalloc(&dbg, 64);
	cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
	cudaMemcpy(db, b_kmajor ? bt.data() : b.data(), b.size() * 4, cudaMemcpyHostToDevice);
	cudaMemset(dc, 0xff, c.size() * 4);
	CUtensorMap ma, mb;
	{
		cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M}; cuuint64_t strides[1] = {K * 4}; cuuint32_t box[2] = {32, 128}; cuuint32_t el[2] = {1, 1};
		CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r) printf("encode A failed %d\n", (int)r);
	}
	{
		cuuint64_t dims[2] = {(cuuint64_t)(b_kmajor ? K : N), (cuuint64_t)(b_kmajor ? N : K)}; cuuint64_t strides[1] = {(cuuint64_t)(b_kmajor ? K : N) * 4};
		cuuint32_t box[2] = {32, (cuuint32_t)(b_kmajor ? 64 : 32)}; cuuint32_t el[2] = {1, 1};
		CUresult r = enc(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r) printf("encode B failed %d\n", (int)r);
	}
	cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
	probe<<<1, 128, 40000>>>(ma, mb, dc, variant, dbg);
	cudaError_t e = cudaDeviceSynchronize();
	cudaMemcpy(c.data(), dc, c.size() * 4, cudaMemcpyDeviceToHost);
	unsigned h[2];
	cudaMemcpy(h, dbg, 8, cudaMemcpyDeviceToHost);
	double maxerr = 0;
	int nz = 0;
	for (int i = 0; i < M * N; i++) {
		maxerr = fmax(maxerr, fabs((double)c[i] - cref[i]));
		nz += c[i] != 0.f;
	}
	printf("variant %d: sync=%s tmem=%08x idesc=%08x nonzero=%d maxerr=%g  c[0..3]=%g %g %g %g  ref=%g %g %g %g  c[row1]=%g ref=%g\n", variant, cudaGetErrorName(e),
	       h[0], h[1], nz, maxerr, c[0], c[1], c[2], c[3], cref[0], cref[1], cref[2], cref[3], c[N], cref[N]);
	return 0;
}
