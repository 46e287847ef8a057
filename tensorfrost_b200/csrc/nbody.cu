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
#ifndef TF_HOST_SIM  // tests/cpu_sim/kernel_on_host.cpp compiles the kernels below for the host through cuda_host_shim.h
#include <cstdlib>

#include "tfcuda_internal.h"
#endif

namespace {

constexpr int NB_THREADS = 128;
constexpr int NB_PER_THREAD = 2;
constexpr int NB_TILE = 256;  // j bodies per shared-memory stage

#ifndef TF_HOST_SIM
__device__ __forceinline__ float fast_rsqrt(float v) {
	float r;
	asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
	return r;
}
#else
static inline float fast_rsqrt(float v) { return 1.0f / sqrtf(v); }  // host stand-in (the hardware instruction is 2 ulp from this)
#endif

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
			else { p.x = p.y = p.z = 0.0f; p.w = 0.0f; }  // padding bodies contribute zero force (tail tile only)
			s_pos[t] = p;
		}
		__syncthreads();
		if (j0 + NB_TILE <= n) {
			// full tile: 3 FADD + 6 FFMA + 2 FMUL + 1 MUFU per interaction.  d2 >= eps > 0 is never denormal, so the raw
			// rsqrt.approx.ftz (no range fix-up around the MUFU) is exact to the same 2 ulp as rsqrtf
#pragma unroll 8
			for (int t = 0; t < NB_TILE; t++) {
				float4 q = s_pos[t];
#pragma unroll
				for (int b = 0; b < NB_PER_THREAD; b++) {
					float dx = px[b] - q.x, dy = py[b] - q.y, dz = pz[b] - q.z;
					float d2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, eps)));
					float inv = fast_rsqrt(d2);
					float w = inv * inv * inv;
					fx[b] = fmaf(-dx, w, fx[b]);
					fy[b] = fmaf(-dy, w, fy[b]);
					fz[b] = fmaf(-dz, w, fz[b]);
				}
			}
		} else {
#pragma unroll 4
			for (int t = 0; t < NB_TILE; t++) {
				float4 q = s_pos[t];
#pragma unroll
				for (int b = 0; b < NB_PER_THREAD; b++) {
					float dx = px[b] - q.x, dy = py[b] - q.y, dz = pz[b] - q.z;
					float d2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, eps)));
					float inv = fast_rsqrt(d2);
					float w = inv * inv * inv * q.w;
					fx[b] = fmaf(-dx, w, fx[b]);
					fy[b] = fmaf(-dy, w, fy[b]);
					fz[b] = fmaf(-dz, w, fz[b]);
				}
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

// ---- packed variant: Blackwell's f32x2 arithmetic (add/mul/fma.f32x2: two fp32 lanes per issued instruction) ------------------
// One thread still owns two i bodies; j bodies are staged SoA so that one LDS.128 brings x (or y, z) of FOUR consecutive j's and
// each f32x2 instruction works on a (j, j+1) pair: per pair and i body 3 sub2 + 3 fma2 + 2 mul2 + 3 fma2 = 11 issue slots + 2 MUFU,
// i.e. 6.5 slots per interaction against 12 for the scalar loop.  Even and odd j's accumulate separately and are added at the end.
typedef unsigned long long u64;
#ifndef TF_HOST_SIM
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#else  // host stand-ins with the same lane layout (low word = first lane) and the same IEEE operations per lane
static inline u64 pack2(float lo, float hi) { return (u64)__float_as_uint(lo) | ((u64)__float_as_uint(hi) << 32); }
static inline void unpack2(u64 v, float& lo, float& hi) { lo = __uint_as_float((unsigned)v); hi = __uint_as_float((unsigned)(v >> 32)); }
static inline u64 sub2(u64 a, u64 b) { float al, ah, bl, bh; unpack2(a, al, ah); unpack2(b, bl, bh); return pack2(al - bl, ah - bh); }
static inline u64 mul2(u64 a, u64 b) { float al, ah, bl, bh; unpack2(a, al, ah); unpack2(b, bl, bh); return pack2(al * bl, ah * bh); }
static inline u64 fma2(u64 a, u64 b, u64 c) {
	float al, ah, bl, bh, cl, ch;
	unpack2(a, al, ah); unpack2(b, bl, bh); unpack2(c, cl, ch);
	return pack2(fmaf(al, bl, cl), fmaf(ah, bh, ch));
}
#endif

constexpr int NX_THREADS = 128;
constexpr int NX_TILE = 512;  // j bodies per stage (SoA: 3 x 2 KB)

__global__ void __launch_bounds__(NX_THREADS) nbody_kernel_x2(const float* __restrict__ x, const float* __restrict__ v, float* __restrict__ x_new,
                                                              float* __restrict__ v_new, int n, float dt, float eps) {
	__shared__ __align__(16) float s_x[NX_TILE], s_y[NX_TILE], s_z[NX_TILE];
	const int first = (blockIdx.x * NX_THREADS + threadIdx.x) * 2;
	float px[2], py[2], pz[2];
	u64 pxx[2], pyy[2], pzz[2], fx[2], fy[2], fz[2];
#pragma unroll
	for (int b = 0; b < 2; b++) {
		int i = min(first + b, n - 1);
		px[b] = x[3 * i + 0]; py[b] = x[3 * i + 1]; pz[b] = x[3 * i + 2];
		pxx[b] = pack2(px[b], px[b]); pyy[b] = pack2(py[b], py[b]); pzz[b] = pack2(pz[b], pz[b]);
		fx[b] = fy[b] = fz[b] = pack2(0.0f, 0.0f);
	}
	const u64 eps2 = pack2(eps, eps);
	// padding bodies of the tail tile sit far away: d2 ~ 3e30, rsqrt^3 flushes to zero, the contribution is exactly 0
	const float kFar = 1.0e15f;
	for (int j0 = 0; j0 < n; j0 += NX_TILE) {
		__syncthreads();
		for (int t = threadIdx.x; t < NX_TILE; t += NX_THREADS) {
			int j = j0 + t;
			bool in = j < n;
			s_x[t] = in ? x[3 * j + 0] : kFar;
			s_y[t] = in ? x[3 * j + 1] : kFar;
			s_z[t] = in ? x[3 * j + 2] : kFar;
		}
		__syncthreads();
#pragma unroll 2
		for (int t = 0; t < NX_TILE; t += 4) {
			const ulonglong2 qx = *reinterpret_cast<const ulonglong2*>(&s_x[t]);
			const ulonglong2 qy = *reinterpret_cast<const ulonglong2*>(&s_y[t]);
			const ulonglong2 qz = *reinterpret_cast<const ulonglong2*>(&s_z[t]);
#pragma unroll
			for (int h = 0; h < 2; h++) {
				const u64 jx = h ? qx.y : qx.x, jy = h ? qy.y : qy.x, jz = h ? qz.y : qz.x;
#pragma unroll
				for (int b = 0; b < 2; b++) {
					u64 ndx = sub2(jx, pxx[b]), ndy = sub2(jy, pyy[b]), ndz = sub2(jz, pzz[b]);  // -(x_i - x_j)
					u64 d2 = fma2(ndx, ndx, fma2(ndy, ndy, fma2(ndz, ndz, eps2)));
					float d2a, d2b;
					unpack2(d2, d2a, d2b);
					u64 inv = pack2(fast_rsqrt(d2a), fast_rsqrt(d2b));
					u64 w = mul2(mul2(inv, inv), inv);
					fx[b] = fma2(ndx, w, fx[b]);
					fy[b] = fma2(ndy, w, fy[b]);
					fz[b] = fma2(ndz, w, fz[b]);
				}
			}
		}
	}
#pragma unroll
	for (int b = 0; b < 2; b++) {
		int i = first + b;
		if (i < n) {
			float ax, bx, ay, by, az, bz;
			unpack2(fx[b], ax, bx); unpack2(fy[b], ay, by); unpack2(fz[b], az, bz);
			float vx = v[3 * i + 0] + (ax + bx) * dt, vy = v[3 * i + 1] + (ay + by) * dt, vz = v[3 * i + 2] + (az + bz) * dt;
			v_new[3 * i + 0] = vx; v_new[3 * i + 1] = vy; v_new[3 * i + 2] = vz;
			x_new[3 * i + 0] = px[b] + vx * dt; x_new[3 * i + 1] = py[b] + vy * dt; x_new[3 * i + 2] = pz[b] + vz * dt;
		}
	}
}

}  // namespace

#ifndef TF_HOST_SIM
extern "C" int tfcuda_nbody_step(uint64_t x, uint64_t v, uint64_t x_new, uint64_t v_new, size_t n, float dt, float eps) {
	tfcuda::State& s = tfcuda::state();
	if (!s.initialized) { tfcuda::set_error("tfcuda_nbody_step: not initialised"); return 1; }
	if (n == 0) return 0;
	if (n > 0x2fffffffull) { tfcuda::set_error("tfcuda_nbody_step: too many bodies"); return 1; }
	tfcuda::ProfileScope prof("lib/nbody");
	unsigned blocks = (unsigned)((n + NB_THREADS * NB_PER_THREAD - 1) / (NB_THREADS * NB_PER_THREAD));
	static const int variant = getenv("TFCUDA_NBODY_VARIANT") ? atoi(getenv("TFCUDA_NBODY_VARIANT")) : 1;  // 1 = packed f32x2 (default), 0 = scalar
	if (variant == 1 && n >= 4 * NX_TILE)
		nbody_kernel_x2<<<(unsigned)((n + NX_THREADS * 2 - 1) / (NX_THREADS * 2)), NX_THREADS, 0, s.stream>>>(
		    reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(v), reinterpret_cast<float*>(x_new), reinterpret_cast<float*>(v_new), (int)n, dt, eps);
	else
		nbody_kernel<<<blocks, NB_THREADS, 0, s.stream>>>(reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(v), reinterpret_cast<float*>(x_new),
		                                                  reinterpret_cast<float*>(v_new), (int)n, dt, eps);
	return tfcuda::check_launch("tfcuda_nbody_step");
}
#endif  // TF_HOST_SIM
