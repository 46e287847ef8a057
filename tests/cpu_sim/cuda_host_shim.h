// TEST INFRASTRUCTURE ONLY (tests/cpu_sim): lets the CUDA C++ that the emitter produces for sm_100a be compiled by g++ and executed on
// the host, one thread at a time, so that the emitter's OUTPUT TEXT (kernel wrapper, binding / variable unpacking, index arithmetic,
// prelude helpers, atomics) can be checked against the reference's golden outputs without a GPU.  Nothing under tensorfrost_b200/
// includes this file; it is not a backend and not a fallback.
//
// Model: blocks run one after the other.  The threads of a block run serially to completion, except for kernels whose text calls
// tf_group_barrier: those get one host thread per CUDA thread of the block and a std::barrier (threads that return early drop out of
// it, as exited CUDA threads do).  __shared__ arrays are statics shared by the threads of the current block; atomics are real atomics.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

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
static thread_local sim_dim3 blockIdx, threadIdx, blockDim, gridDim;
static std::barrier<>* sim_block_barrier = nullptr;  // set by the launcher of a kernel that uses barriers

static inline float __fdiv_rn(float a, float b) { return a / b; }  // IEEE division on the host anyway
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
static inline void __syncthreads() {
	if (sim_block_barrier) sim_block_barrier->arrive_and_wait();
}

static inline unsigned atomicCAS(unsigned* p, unsigned expected, unsigned desired) {
	__atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
	return expected;  // the value seen
}
template <typename T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <> inline float atomicAdd<float>(float* p, float v) {
	unsigned* u = reinterpret_cast<unsigned*>(p);
	unsigned cur = __atomic_load_n(u, __ATOMIC_SEQ_CST);
	for (;;) {
		unsigned want = __float_as_uint(__uint_as_float(cur) + v);
		unsigned seen = atomicCAS(u, cur, want);
		if (seen == cur) return __uint_as_float(cur);
		cur = seen;
	}
}
template <typename T> static inline T sim_atomic_rmw(T* p, T v, T (*op)(T, T)) {
	T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
	while (!__atomic_compare_exchange_n(p, &cur, op(cur, v), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
	return cur;
}
template <typename T> static inline T atomicMin(T* p, T v) { return sim_atomic_rmw<T>(p, v, [](T a, T b) { return b < a ? b : a; }); }
template <typename T> static inline T atomicMax(T* p, T v) { return sim_atomic_rmw<T>(p, v, [](T a, T b) { return b > a ? b : a; }); }
template <typename T> static inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <typename T> static inline T atomicXor(T* p, T v) { return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST); }
