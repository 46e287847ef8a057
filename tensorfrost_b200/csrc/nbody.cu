// One all-pairs gravity step.  X, V: [N,3] fp32.
//
// The reference program (examples/Simulation/n-body-benchmark.py:16-34) compiles to ONE fused kernel of
// shape [N,3]: a thread per (body, component) loops over every j, recomputing the full distance three
// times and calling pow(x, 2.f).  This kernel computes the same step
//     dx = X[i]-X[j];  d2 = |dx|^2 + eps;  F_i = sum_j -dx / (d2*sqrt(d2));  V' = V + F dt;  X' = X + V' dt
// with one thread per TWO bodies, the j bodies staged through shared memory in tiles (float4, one
// LDS.128 broadcast per j), rsqrt on the SFU and FMAs elsewhere.  It is compute bound on the FP32/SFU pipes
// (X is 3 MB at N=262144 and stays in L2), not HBM and not tensor cores.
// The sum over j is per-thread serial in j order like the oracle's loop; -dx/(d2*sqrt(d2)) is evaluated as
// -dx * rsqrt(d2)^3, a few ulp from the oracle's divide (tests bound the step at 1e-5 relative).
#include "tfcuda_internal.h"

namespace {

constexpr int NB_THREADS = 128;
constexpr int NB_PER_THREAD = 2;
constexpr int NB_TILE = 256;  // j bodies per shared-memory stage

__global__ void __launch_bounds__(NB_THREADS) nbody_kernel(const float* __restrict__ x, const float* __restrict__ v, float* __restrict__ x_new,
                                                           float* __restrict__ v_new, int n, float dt, float eps) {
	__shared__ float4 s_pos[NB_TILE];
	const int first = (blockIdx.x * NB_THREADS + threadIdx.x) * NB_PER_THREAD;
	float px[NB_PER_THREAD], py[NB_PER_THREAD], pz[NB_PER_THREAD];
	float fx[NB_PER_THREAD], fy[NB_PER_THREAD], fz[NB_PER_THREAD];
#pragma unroll
	for (int b = 0; b < NB_PER_THREAD; b++) {
		int i = min(first + b, n - 1);
		px[b] = x[3 * i + 0]; py[b] = x[3 * i + 1]; pz[b] = x[3 * i + 2];
		fx[b] = fy[b] = fz[b] = 0.0f;
	}
	for (int j0 = 0; j0 < n; j0 += NB_TILE) {
		__syncthreads();
		for (int t = threadIdx.x; t < NB_TILE; t += NB_THREADS) {
			int j = j0 + t;
			float4 p;
			if (j < n) { p.x = x[3 * j + 0]; p.y = x[3 * j + 1]; p.z = x[3 * j + 2]; p.w = 1.0f; }
			else { p.x = p.y = p.z = 0.0f; p.w = 0.0f; }  // padding bodies contribute zero force
			s_pos[t] = p;
		}
		__syncthreads();
#pragma unroll 8
		for (int t = 0; t < NB_TILE; t++) {
			float4 q = s_pos[t];
#pragma unroll
			for (int b = 0; b < NB_PER_THREAD; b++) {
				float dx = px[b] - q.x, dy = py[b] - q.y, dz = pz[b] - q.z;
				float d2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, eps)));
				float inv = rsqrtf(d2);
				float w = -(inv * inv * inv) * q.w;
				fx[b] = fmaf(dx, w, fx[b]);
				fy[b] = fmaf(dy, w, fy[b]);
				fz[b] = fmaf(dz, w, fz[b]);
			}
		}
	}
#pragma unroll
	for (int b = 0; b < NB_PER_THREAD; b++) {
		int i = first + b;
		if (i < n) {
			float vx = v[3 * i + 0] + fx[b] * dt, vy = v[3 * i + 1] + fy[b] * dt, vz = v[3 * i + 2] + fz[b] * dt;
			v_new[3 * i + 0] = vx; v_new[3 * i + 1] = vy; v_new[3 * i + 2] = vz;
			x_new[3 * i + 0] = px[b] + vx * dt; x_new[3 * i + 1] = py[b] + vy * dt; x_new[3 * i + 2] = pz[b] + vz * dt;
		}
	}
}

}  // namespace

extern "C" int tfcuda_nbody_step(uint64_t x, uint64_t v, uint64_t x_new, uint64_t v_new, size_t n, float dt, float eps) {
	tfcuda::State& s = tfcuda::state();
	if (!s.initialized) { tfcuda::set_error("tfcuda_nbody_step: not initialised"); return 1; }
	if (n == 0) return 0;
	if (n > 0x2fffffffull) { tfcuda::set_error("tfcuda_nbody_step: too many bodies"); return 1; }
	tfcuda::ProfileScope prof("lib/nbody");
	unsigned blocks = (unsigned)((n + NB_THREADS * NB_PER_THREAD - 1) / (NB_THREADS * NB_PER_THREAD));
	nbody_kernel<<<blocks, NB_THREADS, 0, s.stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(v), reinterpret_cast<float*>(x_new),
	                                                  reinterpret_cast<float*>(v_new), (int)n, dt, eps);
	return tfcuda::check_launch("tfcuda_nbody_step");
}
