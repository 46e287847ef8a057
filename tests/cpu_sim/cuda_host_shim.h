// TEST INFRASTRUCTURE ONLY (tests/cpu_sim): lets the CUDA C++ that the emitter produces for sm_100a be compiled by g++ and executed on
// the host, one thread at a time, so that the emitter's OUTPUT TEXT (kernel wrapper, binding / variable unpacking, index arithmetic,
// prelude helpers, atomics) can be checked against the reference's golden outputs without a GPU.  Nothing under tensorfrost_b200/
// includes this file; it is not a backend and not a fallback.
//
// Model: blocks and the threads of a block run serially to completion (so atomics are plain read-modify-writes and __shared__ arrays
// are statics shared by the serial threads of the current block).  Kernels that need a real barrier cannot run this way: the runner
// refuses any kernel whose text calls tf_group_barrier.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define TF_HOST_SIM 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __restrict__
#define __shared__ static

struct sim_dim3 { unsigned x, y, z; };
static sim_dim3 blockIdx, threadIdx, blockDim, gridDim;

static inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
static inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
static inline unsigned __brev(unsigned v) {
	v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
	v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
	v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
	v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
	return (v >> 16) | (v << 16);
}
static inline void __syncthreads() {}

template <typename T> static inline T atomicAdd(T* p, T v) { T old = *p; *p = old + v; return old; }
template <typename T> static inline T atomicMin(T* p, T v) { T old = *p; *p = v < old ? v : old; return old; }
template <typename T> static inline T atomicMax(T* p, T v) { T old = *p; *p = v > old ? v : old; return old; }
template <typename T> static inline T atomicAnd(T* p, T v) { T old = *p; *p = old & v; return old; }
template <typename T> static inline T atomicOr(T* p, T v) { T old = *p; *p = old | v; return old; }
template <typename T> static inline T atomicXor(T* p, T v) { T old = *p; *p = old ^ v; return old; }
static inline unsigned atomicCAS(unsigned* p, unsigned expected, unsigned desired) {
	unsigned old = *p;
	if (old == expected) *p = desired;
	return old;
}
