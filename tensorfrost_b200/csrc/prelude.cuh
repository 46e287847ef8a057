// Device prelude for kernels emitted by the CUDA code generator (tensorfrost_b200/overlay/.../CUDA.cpp).
// It is the sm_100a restatement of the reference's C++ helper header
// (TensorFrost/Backend/CodeGen/Langs/CPP.cpp:31-271): every helper an emitted kernel can call, with the
// SAME arithmetic as the oracle so fp32 results agree to the last bit wherever libm/CUDA agree.
// Compiled by NVRTC (no system headers), so only builtins are used.
//
// Naming: the emitter maps every IR function op `f` to `tf_f` (name_map_), so user variable names can
// never shadow a helper.  Buffers are arrays of 32-bit words (`uint`), typed access is by bit
// reinterpretation (Generators.cpp:399-437).

typedef unsigned int uint;

#ifndef TF_QUIRK_INT_AND_IS_OR
// The oracle implements InterlockedAnd(int*) with fetch_or (CPP.cpp:183-187).  Parity with the
// C++/OpenMP backend means reproducing it; build kernels with -DTF_QUIRK_INT_AND_IS_OR=0 for a true AND.
#define TF_QUIRK_INT_AND_IS_OR 1
#endif

#define TF_DEV static __device__ __forceinline__

// read-only buffer bindings; build kernels with -DTF_NO_RESTRICT when a program binds one buffer both read-only and writable
#ifdef TF_NO_RESTRICT
#define TF_RO const uint*
#else
#define TF_RO const uint* __restrict__
#endif

// ---- programmatic dependent launch (EXPERIMENTAL, TFCUDA_PDL=1 at trace time AND at run time; not yet run on hardware) ----------
// First statement of every emitted kernel when enabled: let the next kernel of the stream start launching its CTAs as soon as all of
// ours are resident, then wait until every kernel before us has completed and flushed its memory.  Nothing is read or written before
// the wait, so the program order of the stream is preserved; what overlaps is the launch latency of short dependent kernels
// (the 32 multigrid sweeps of the fluid step: 3.6 us each for 3 MB of L2-resident data).
#ifndef TF_HOST_SIM  // tests/cpu_sim runs the emitted text on the host, where this PTX does not exist
TF_DEV void tf_pdl_prologue() { asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory"); }
#else
TF_DEV void tf_pdl_prologue() {}
#endif

// ---- value-range facts the emitter states about its own index arithmetic ------------------------------------------------------
// `TF_ASSUME(block_id >= 0)` opens every emitted kernel: a dispatch never has more than 2^31-1 blocks (tfcuda_launch refuses larger ones),
// so the block id, and every index_k = block * group + thread derived from it, is non-negative.  The compiler cannot know that from
// `(int)(blockIdx.x + offset)`; told so, and with `index_k < extent` known inside the dispatch guard, it deletes the lower half of every
// clamp on an index, the `i >= 0` halves of the generated out-of-bounds tests and the sign fix-ups of the constant divisions
// (SASS, round 2: fluid multigrid sweeps 81 -> 66 instructions per thread, NCA filter-backward kernels 100 -> 75).
#ifdef TF_HOST_SIM  // tests/cpu_sim compiles this text with g++, which has no __builtin_assume
#define TF_ASSUME(x) ((void)0)
#else
#define TF_ASSUME(x) __builtin_assume(x)
#endif

// ---- bit casts (CPP.cpp:53-96) ------------------------------------------------------------
TF_DEV float asfloat(uint x) { return __uint_as_float(x); }
TF_DEV float asfloat(int x) { return __int_as_float(x); }
TF_DEV float asfloat(float x) { return x; }
TF_DEV uint asuint(float x) { return __float_as_uint(x); }
TF_DEV uint asuint(int x) { return (uint)x; }
TF_DEV uint asuint(uint x) { return x; }
TF_DEV uint asuint(bool x) { return x ? 1u : 0u; }
TF_DEV int asint(uint x) { return (int)x; }
TF_DEV int asint(int x) { return x; }
TF_DEV int asint(float x) { return __float_as_int(x); }
TF_DEV int asint(bool x) { return x ? 1 : 0; }
TF_DEV bool asbool(uint x) { return x != 0u; }
TF_DEV bool asbool(int x) { return x != 0; }
TF_DEV bool asbool(bool x) { return x; }
TF_DEV bool asbool(float x) { return __float_as_uint(x) != 0u; }

// ---- min / max / clamp: `a < b ? a : b`, NOT fminf (NaN behaviour follows the oracle, CPP.cpp:33-51,98-106)
TF_DEV int tf_min(int a, int b) { return a < b ? a : b; }
TF_DEV int tf_max(int a, int b) { return a > b ? a : b; }
TF_DEV uint tf_min(uint a, uint b) { return a < b ? a : b; }
TF_DEV uint tf_max(uint a, uint b) { return a > b ? a : b; }
TF_DEV float tf_min(float a, float b) { return a < b ? a : b; }
TF_DEV float tf_max(float a, float b) { return a > b ? a : b; }
TF_DEV int tf_clamp(int x, int a, int b) { return tf_min(tf_max(x, a), b); }
TF_DEV uint tf_clamp(uint x, uint a, uint b) { return tf_min(tf_max(x, a), b); }
TF_DEV float tf_clamp(float x, float a, float b) { return tf_min(tf_max(x, a), b); }

// ---- small math helpers (CPP.cpp:108-139) ---------------------------------------------------
TF_DEV float tf_lerp(float a, float b, float t) { return a + (b - a) * t; }
TF_DEV float tf_smoothstep(float a, float b, float t) {
	t = tf_clamp((t - a) / (b - a), 0.0f, 1.0f);
	return t * t * (3.0f - 2.0f * t);
}
TF_DEV float tf_sign(float x) { return x < 0.0f ? -1.0f : 1.0f; }  // sign(0) == +1, as the oracle
TF_DEV int tf_sign(int x) { return x < 0 ? -1 : 1; }
TF_DEV uint tf_reversebits(uint x) { return __brev(x); }
TF_DEV int tf_reversebits(int x) { return (int)__brev((uint)x); }
TF_DEV float tf_abs(float x) { return fabsf(x); }
TF_DEV int tf_abs(int x) { return x < 0 ? -x : x; }
TF_DEV uint tf_abs(uint x) { return x; }

// libm-backed ops: the oracle falls through to <cmath> float overloads (CPP.cpp:11 renames sqrt only)
TF_DEV float tf_ceil(float x) { return ceilf(x); }
TF_DEV float tf_floor(float x) { return floorf(x); }
TF_DEV float tf_round(float x) { return roundf(x); }  // half away from zero, like C round()
TF_DEV float tf_trunc(float x) { return truncf(x); }
TF_DEV float tf_exp(float x) { return expf(x); }
TF_DEV float tf_exp2(float x) { return exp2f(x); }
TF_DEV float tf_log(float x) { return logf(x); }
TF_DEV float tf_log2(float x) { return log2f(x); }
TF_DEV float tf_sqrt(float x) { return sqrtf(x); }
TF_DEV float tf_sin(float x) { return sinf(x); }
TF_DEV float tf_cos(float x) { return cosf(x); }
TF_DEV float tf_tan(float x) { return tanf(x); }
TF_DEV float tf_asin(float x) { return asinf(x); }
TF_DEV float tf_acos(float x) { return acosf(x); }
TF_DEV float tf_atan(float x) { return atanf(x); }
TF_DEV float tf_sinh(float x) { return sinhf(x); }
TF_DEV float tf_cosh(float x) { return coshf(x); }
TF_DEV float tf_tanh(float x) { return tanhf(x); }
// x ** 2.0 is how user programs square (n-body, losses): when the exponent is the literal 2 the branch folds at compile time to one
// multiply - the correctly rounded square, which is also what the oracle's libm returns - instead of a ~40-instruction powf
TF_DEV float tf_pow(float a, float b) { return b == 2.0f ? a * a : powf(a, b); }
TF_DEV float tf_atan2(float a, float b) { return atan2f(a, b); }
TF_DEV float tf_fma(float a, float b, float c) { return fmaf(a, b, c); }
// Ops the op table lists (Operations.cpp:150-173) but the C++ helper header never defines, so they do
// not compile on the oracle ("parity unpinned"); defined here with their HLSL/GLSL meaning.
TF_DEV float tf_frac(float x) { return x - floorf(x); }
TF_DEV float tf_rcp(float x) { return 1.0f / x; }
TF_DEV float tf_rsqrt(float x) { return 1.0f / sqrtf(x); }
TF_DEV float tf_step(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
TF_DEV float tf_modf(float a, float b) { return a - b * floorf(a / b); }  // GLSL mod (GLSL.cpp:11)

// ---- pcg hash (CPP.cpp:261-271) --------------------------------------------------------------
TF_DEV uint tf_pcg(uint v) {
	uint state = v * 747796405u + 2891336453u;
	uint word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
	return (word >> 22u) ^ word;
}
// __fdiv_rn: correctly rounded whatever division mode the kernels are compiled with - random streams (dropout-style masks, NCA's
// fire rate) must be the reference's bit for bit
TF_DEV float tf_pcgf(uint v) { return __fdiv_rn((float)tf_pcg(v), (float)0xffffffffu); }

// ---- barrier: a real one (the oracle's is a no-op, CPP.cpp:259) -------------------------------
TF_DEV void tf_group_barrier() { __syncthreads(); }

// ---- atomics on word buffers (CPP.cpp:141-257).  `mem` is the uint buffer (global or shared),
// reinterpret per element type.  Float add is the native red/atom.add.f32, not a CAS loop. --------
//
// Warp-aggregated add (tf.scatterAdd, the autodiff of `load`: Implementations.cpp:185-194; broadcast gradients): the lanes of a warp
// that are about to add to the SAME word are found with match.any, combined in registers (lane order: deterministic inside the warp)
// and the group's lowest lane issues ONE red.global.add.  A lane alone at its address goes straight to the atomic.
// OFF by default, measured on the B200 (round 2, NCA training step, batch 256): 838 ms per iteration with aggregation, 564 ms with
// plain red.global.add.f32 - the scatter-adds of that program mostly hit distinct addresses inside a warp, so match.any (plus the
// divergent combine loop) costs far more than the few atomics it saves, and L2 resolves the remaining conflicts at line rate.
// Build kernels with -DTF_WARP_AGG_ATOMICS=1 for programs whose warps really collide (e.g. histograms into a few bins); the
// library kernel tfcuda_scatter_add (csrc/scatter.cu) always aggregates.
#ifndef TF_WARP_AGG_ATOMICS
#define TF_WARP_AGG_ATOMICS 0
#endif
#if TF_WARP_AGG_ATOMICS && !defined(TF_HOST_SIM)
TF_DEV uint tf_lane_id() {
	uint l;
	asm("mov.u32 %0, %%laneid;" : "=r"(l));
	return l;
}
TF_DEV void tf_atomic_add(uint* mem, int a, float v) {
	const uint active = __activemask();
	const uint peers = __match_any_sync(active, (unsigned long long)(mem + a));
	const int members = __popc(peers);
	if (members == 1) {
		atomicAdd((float*)mem + a, v);
		return;
	}
	float sum = 0.0f;
	for (int k = 1; k <= members; k++) sum += __shfl_sync(peers, v, (int)__fns(peers, 0, k));
	if (tf_lane_id() == (uint)(__ffs((int)peers) - 1)) atomicAdd((float*)mem + a, sum);
}
TF_DEV void tf_atomic_add(uint* mem, int a, uint v) {
	const uint peers = __match_any_sync(__activemask(), (unsigned long long)(mem + a));
	if (__popc(peers) > 1) v = __reduce_add_sync(peers, v);
	if (tf_lane_id() == (uint)(__ffs((int)peers) - 1)) atomicAdd(mem + a, v);
}
TF_DEV void tf_atomic_add(uint* mem, int a, int v) {
	const uint peers = __match_any_sync(__activemask(), (unsigned long long)(mem + a));
	if (__popc(peers) > 1) v = __reduce_add_sync(peers, v);
	if (tf_lane_id() == (uint)(__ffs((int)peers) - 1)) atomicAdd((int*)mem + a, v);
}
#else
TF_DEV void tf_atomic_add(uint* mem, int a, uint v) { atomicAdd(mem + a, v); }
TF_DEV void tf_atomic_add(uint* mem, int a, int v) { atomicAdd((int*)mem + a, v); }
// A float add of exactly zero is skipped: it cannot change the sum (at most the sign of a zero), and the reference's autodiff produces
// such adds in bulk - the gradient of a concatenation scatters `in_this_half ? g : 0.f` into BOTH halves, with the index of the wrong half
// clamped to one edge element, so a third of NCA's split kernel's atomics were zeros queueing on the same few addresses.  NaN is not
// skipped (NaN != 0).  -DTF_ATOMIC_ADD_ZERO=1 restores the unconditional add.
#ifndef TF_ATOMIC_ADD_ZERO
#define TF_ATOMIC_ADD_ZERO 0
#endif
TF_DEV void tf_atomic_add(uint* mem, int a, float v) {
	if (TF_ATOMIC_ADD_ZERO || v != 0.0f) atomicAdd((float*)mem + a, v);
}
#endif
TF_DEV uint tf_atomic_add_prev(uint* mem, int a, uint v) { return atomicAdd(mem + a, v); }
TF_DEV int tf_atomic_add_prev(uint* mem, int a, int v) { return atomicAdd((int*)mem + a, v); }
TF_DEV float tf_atomic_add_prev(uint* mem, int a, float v) { return atomicAdd((float*)mem + a, v); }
TF_DEV void tf_atomic_min(uint* mem, int a, uint v) { atomicMin(mem + a, v); }
TF_DEV void tf_atomic_min(uint* mem, int a, int v) { atomicMin((int*)mem + a, v); }
TF_DEV void tf_atomic_max(uint* mem, int a, uint v) { atomicMax(mem + a, v); }
TF_DEV void tf_atomic_max(uint* mem, int a, int v) { atomicMax((int*)mem + a, v); }
TF_DEV void tf_atomic_min(uint* mem, int a, float v) {
	uint* p = mem + a;
	uint cur = *p;
	for (;;) {
		float goal = tf_min(__uint_as_float(cur), v);
		uint seen = atomicCAS(p, cur, __float_as_uint(goal));
		if (seen == cur) break;
		cur = seen;
	}
}
TF_DEV void tf_atomic_max(uint* mem, int a, float v) {
	uint* p = mem + a;
	uint cur = *p;
	for (;;) {
		float goal = tf_max(__uint_as_float(cur), v);
		uint seen = atomicCAS(p, cur, __float_as_uint(goal));
		if (seen == cur) break;
		cur = seen;
	}
}
TF_DEV void tf_atomic_and(uint* mem, int a, uint v) { atomicAnd(mem + a, v); }
TF_DEV void tf_atomic_and(uint* mem, int a, int v) {
#if TF_QUIRK_INT_AND_IS_OR
	atomicOr((int*)mem + a, v);
#else
	atomicAnd((int*)mem + a, v);
#endif
}
TF_DEV void tf_atomic_or(uint* mem, int a, uint v) { atomicOr(mem + a, v); }
TF_DEV void tf_atomic_or(uint* mem, int a, int v) { atomicOr((int*)mem + a, v); }
TF_DEV void tf_atomic_xor(uint* mem, int a, uint v) { atomicXor(mem + a, v); }
TF_DEV void tf_atomic_xor(uint* mem, int a, int v) { atomicXor((int*)mem + a, v); }

// `discard` keyword op (Operations.cpp:33): end this thread.
#define discard return
