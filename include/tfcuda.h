/*
 * tfcuda.h — C-ABI of libtfcuda.so, the B200 (sm_100a) execution backend for TensorFrost programs.
 *
 * This is the drop-in boundary.  A compiled TensorFrost program is a host function
 *     extern "C" int main(TFTensor* in, TFTensor* out, TFRuntime runtime)
 * (reference: TensorFrost/Backend/CodeGen/Langs/CPP.cpp:644-674, typedef main_func in
 * TensorFrost/Backend/TensorMemory.h:71) that talks to its backend only through the
 * TFRuntime callback table.  libtfcuda.so provides that table (tfcuda_runtime), the device
 * buffers behind TFBuffer, the NVRTC/driver-API kernel registry behind TFDispatchInfo.kernel_id,
 * and the hand-written sm_100a library kernels (reduce / scan / radix sort / scatter-add /
 * matmul / n-body) that replace the reference's generic per-thread serial lowering.
 *
 * Plain C, pointers and sizes only; no torch / pybind types.  Every entry point names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Error model: functions returning int give 0 on success and a non-zero code on failure with the
 * message in tfcuda_last_error().  The six TFRuntime callbacks instead THROW std::runtime_error,
 * because that is the reference's contract (generated host code and ExecuteProgram propagate C++
 * exceptions: Backend/Backend.cpp:96-107,166-170; CPP.cpp:417,436-446).
 * There is no CPU fallback anywhere: without a CUDA device tfcuda_init fails and nothing runs.
 */
#ifndef TFCUDA_H_
#define TFCUDA_H_

#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * ABI structs.  Bit-for-bit mirrors of the reference structs; when this header is included from
 * a translation unit that already has the reference's definitions (the in-tree backend glue),
 * define TFCUDA_NO_ABI_STRUCTS.
 *   TFType / TFDataFormat : TensorFrost/Compiler/Operations.h:18-55
 *   TFBuffer..TFRuntime   : TensorFrost/Backend/TensorMemory.h:19-72 (repeated in every generated
 *                           host library, CPP.cpp:273-362)
 * ------------------------------------------------------------------------------------------ */
#ifndef TFCUDA_NO_ABI_STRUCTS
typedef enum TFType { TFFloat = 0, TFUint = 1, TFInt = 2, TFBool = 3, TFNone = 4 } TFType;

typedef struct TFDataFormat {
	TFType type;
	size_t size; /* bits per element; always 32 today */
} TFDataFormat;

typedef struct TFBuffer {
	size_t size;            /* capacity in 32-bit WORDS */
	size_t used_size;       /* words in use by the current tensor */
	size_t time_since_used; /* pool ticks since released */
	bool up_to_date;
	bool read_only;
	const char* name;
} TFBuffer;

typedef struct TFTensor {
	TFBuffer* buffer;
	TFDataFormat format;
	size_t dim;
	const size_t* shape; /* Python order: outermost first */
} TFTensor;

typedef struct TFDispatchInfo {
	size_t kernel_id;                  /* process-global id (Backend/KernelManager.cpp:4-8) */
	size_t read_write_count;           /* rw tensors followed by ro tensors, binding order */
	const TFTensor* read_write_tensors;
	size_t read_only_count;            /* always 0: merged upstream (CPP.cpp:485-491) */
	const TFTensor* read_only_tensors;
	size_t variable_count;             /* kernel scalars as raw words + trailing block offset */
	const uint32_t* variables;
	size_t work_group_count;           /* 1-D grid size (CPP.cpp:503-515) */
} TFDispatchInfo;

typedef TFTensor tf_alloc_func(const char* name, const size_t* shape, size_t dim, TFDataFormat fmt, void* user);
typedef void tf_dealloc_func(TFTensor t, void* user);
typedef uint32_t tf_readback_func(TFTensor t, size_t word_index, void* user);
typedef void tf_writeback_func(TFTensor t, size_t word_index, uint32_t value, void* user);
typedef void tf_dispatch_func(TFDispatchInfo info, void* user);
typedef void tf_region_func(const char* name, bool begin, void* user);

typedef struct TFRuntime {
	tf_alloc_func* alloc;
	tf_dealloc_func* dealloc;
	tf_readback_func* readback;
	tf_writeback_func* writeback;
	tf_dispatch_func* dispatch;
	tf_region_func* region;
	void* custom_data;
} TFRuntime;
#endif /* TFCUDA_NO_ABI_STRUCTS */

/* ------------------------------------------------------------------------------------------
 * Lifecycle.  Replaces the backend switch of InitializeBackend (Backend/Backend.cpp:10-73).
 * One device, one stream, one process (the reference's backend is a process-global singleton).
 * ------------------------------------------------------------------------------------------ */
int tfcuda_init(int device);            /* idempotent; fails (non-zero) when no CUDA device */
int tfcuda_is_initialized(void);
int tfcuda_shutdown(void);
const char* tfcuda_last_error(void);
int tfcuda_device_sm_count(void);
int tfcuda_device_index(void);          /* CUDA ordinal the backend runs on (-1 before tfcuda_init) */
const char* tfcuda_device_name(void);
void* tfcuda_stream(void);              /* the CUstream every launch / copy is ordered on */
int tfcuda_sync(void);                  /* cuStreamSynchronize on that stream */

/* The callback table a generated host program is called with.
 * Replaces {Allocator,Deallocator,Readback,Writeback,Dispatch,Region} (Backend/Backend.cpp:96-133)
 * as passed at Backend/Backend.cpp:167. */
TFRuntime tfcuda_runtime(void);

/* ------------------------------------------------------------------------------------------
 * Device buffers.  Replace CpuMemoryManager::CreateBuffer/DeleteBuffer and
 * TFCPUBuffer::{SetDataAtOffset,GetDataAtOffset} (Backend/Backends/CPU/Memory.h:18-59),
 * i.e. the virtuals of TensorMemoryManager / TFBufferTemplate (Backend/TensorMemory.h:74-125).
 * Sizes and offsets are in 32-bit words, as in the reference.
 * ------------------------------------------------------------------------------------------ */
TFBuffer* tfcuda_buffer_create(size_t words);
void tfcuda_buffer_destroy(TFBuffer* buffer);
uint64_t tfcuda_buffer_device_ptr(const TFBuffer* buffer);
int tfcuda_buffer_write(TFBuffer* buffer, size_t word_offset, const uint32_t* src, size_t words);
int tfcuda_buffer_read(const TFBuffer* buffer, size_t word_offset, uint32_t* dst, size_t words);
/* raw device-pointer variants (used by tests / bench and by the in-tree glue) */
int tfcuda_memcpy_h2d(uint64_t dst, const void* src, size_t bytes);
int tfcuda_memcpy_d2h(void* dst, uint64_t src, size_t bytes);
int tfcuda_memcpy_d2d(uint64_t dst, uint64_t src, size_t bytes);
/* Copy engines (the fast path of PyTensorMemory upload/readback, Frontend/Python/PyTensorMemory.cpp:15-84, for pipelines): uploads
 * and downloads run on their own streams so host->device, kernels and device->host overlap.  h2d_async starts once the work queued
 * on the runtime stream so far is done (the previous users of dst); kernels queued after tfcuda_wait_uploads see the data.
 * d2h_async starts once the work queued so far is done; the host may read dst after tfcuda_copy_sync.  Host memory should be
 * page-locked (tfcuda_host_alloc); source / destination device memory must stay allocated until tfcuda_copy_sync. */
int tfcuda_memcpy_h2d_async(uint64_t dst, const void* src, size_t bytes);
int tfcuda_wait_uploads(void);
int tfcuda_memcpy_d2h_async(void* dst, uint64_t src, size_t bytes);
int tfcuda_copy_sync(void);
/* downloads started / completed so far (a caller that owns the source buffers releases them once done >= its ticket) */
uint64_t tfcuda_downloads_issued(void);
uint64_t tfcuda_downloads_done(void);
int tfcuda_memset32(uint64_t dst, uint32_t value, size_t words);
uint64_t tfcuda_malloc(size_t bytes);   /* 0 on failure */
int tfcuda_free(uint64_t ptr);
/* pool statistics: TensorMemoryManager::GetAllocatedSize / GetUnusedAllocatedSize
 * (Backend/TensorMemory.cpp:84-103), in words */
size_t tfcuda_pool_allocated_words(void);
size_t tfcuda_pool_unused_words(void);
/* cudaMallocAsync + cudaFreeAsync calls made for tensor buffers since initialisation (deleted buffers are parked by size and
 * reused without a driver call: Backend/TensorMemory.cpp:145-160 deletes and re-creates buffers every step otherwise) */
uint64_t tfcuda_pool_driver_calls(void);

/* ------------------------------------------------------------------------------------------
 * Kernels.  Replaces CompileKernels (Backend/Backend.cpp:75-94) +
 * OpenGLKernelManager::{CompileKernel,DispatchKernel} (Backend/Backends/OpenGL/KernelManager.h:
 * 73-159) + CpuKernelManager::DispatchKernel (Backend/Backends/CPU/KernelManager.h:28-38).
 *
 * Emitted kernel contract (what the CUDA emitter produces, see tfcuda_prelude):
 *   struct kernel_<id>_args { uint* mem[n_mem]; uint var[n_var]; };
 *   extern "C" __global__ void kernel_<id>(const __grid_constant__ kernel_<id>_args a);
 * n_var counts the trailing _kernel_block_offset word.  Grid = work_group_count x 1 x 1,
 * block = group[0] x group[1] x group[2] (group sizes are baked into the kernel:
 * kernel->root->group_size, CPP.cpp:622-629).  A kernel the emitter coarsened (several "lanes" per
 * thread) is launched with a SMALLER block than the group the host program's block count was
 * computed for; its text states the launch block as a comment `// tfcuda_block: x y z`, and that
 * is what `group` below must carry.  At most 2^31-1 block ids per dispatch (they are `int` in the
 * generated code, and emitted kernels tell the compiler so).
 * ------------------------------------------------------------------------------------------ */
typedef struct TFCudaKernelSource {
	size_t kernel_id;     /* TFDispatchInfo.kernel_id this source serves */
	const char* entry;    /* extern "C" symbol, e.g. "kernel_12" */
	const char* source;   /* CUDA C++ text of this kernel WITHOUT the prelude */
	unsigned group[3];    /* threads per block the kernel is LAUNCHED with, innermost first */
	unsigned n_mem;       /* number of buffer bindings */
	unsigned n_var;       /* number of 32-bit scalar words incl. _kernel_block_offset */
	unsigned library_op;  /* 0 = emitted source; otherwise a TFCUDA_LIB_* id (source may be "") */
} TFCudaKernelSource;

/* The device prelude every emitted kernel is compiled against: the CUDA restatement of the
 * reference's C++ helper header (CPP.cpp:31-271: min/max/clamp/lerp/sign/reversebits, the as-type bit casts,
 * the Interlocked atomics, pcg / pcgf and group_barrier). */
const char* tfcuda_prelude(void);

/* Compile prelude + source for sm_100a WITHOUT a device (no module is loaded): returns 0 when NVRTC accepts it,
 * otherwise the log is in tfcuda_last_error().  Lets emitted kernels be validated on a CPU-only build box. */
int tfcuda_nvrtc_check(const char* source, const char* options);

/* Compile a batch of kernels with NVRTC for sm_100a (chunked, multi-threaded, cubins cached on
 * disk by source hash) and register them under their kernel ids.  options: extra NVRTC flags,
 * space separated (the reference's kernel_compile_options string, PybindModule.cpp:112-115). */
int tfcuda_compile_kernels(const TFCudaKernelSource* kernels, size_t count, const char* options);
/* Directory of the on-disk compile cache (cubins; the backend glue keeps compiled host programs there too, replacing the
 * fixed /tmp/generated_lib_<id>.cpp of Backends/CPU/KernelCompiler.cpp:96-113).  $TFCUDA_CACHE_DIR, else $XDG_CACHE_HOME/tfcuda,
 * else ~/.cache/tfcuda; must be owned by the calling user with mode 0700.  "" when the cache is disabled (TFCUDA_NO_CACHE set, or
 * no safe directory). */
const char* tfcuda_cache_dir(void);

/* Launch by raw device pointers (mem[i] = device address of binding i). */
int tfcuda_launch(size_t kernel_id, const uint64_t* mem, size_t n_mem,
                  const uint32_t* vars, size_t n_var, size_t work_group_count);
/* Graph replay of a program's launches (whole-program CUDA graph).  The backend glue brackets every program execution
 * (ExecuteProgram, Backend/Backend.cpp:137-185) with begin/end; in between tfcuda_launch only records, and the recorded chain is
 * issued as ONE cudaGraphLaunch when the program ends or the moment anything else needs the stream (tf.read, a copy, a library
 * kernel).  Executable graphs are cached by kernel sequence + argument bytes and built the second time a chain is seen with the same
 * arguments, so a steady-state step costs one graph launch and chains that never repeat are simply launched.
 * Bit-identical to eager execution; TFCUDA_GRAPH=0 disables it.  Calls nest; end returns non-zero when a deferred launch failed. */
typedef struct TFCudaGraphStats {
	int enabled;
	uint64_t replays;         /* graph launches */
	uint64_t exact_hits;      /* ... that reused an executable graph unchanged */
	uint64_t patched;         /* (unused since the second-sight policy; kept for ABI stability) */
	uint64_t instantiated;    /* ... that needed a new executable graph */
	uint64_t eager_launches;  /* kernels of short or never-repeating chains launched one by one */
	double host_us;           /* host time spent issuing recorded chains so far (graph upkeep + launches) */
} TFCudaGraphStats;
int tfcuda_graph_begin(void);
int tfcuda_graph_end(void);
int tfcuda_graph_stats(TFCudaGraphStats* out);

/* Launch from the reference's dispatch record (buffers must come from tfcuda_buffer_create). */
int tfcuda_dispatch(const TFDispatchInfo* info);
/* Counters: kernels launched since init (emitted + library), for bench.py's gpu_launches. */
uint64_t tfcuda_launch_count(void);
/* CUDA-event timing on the runtime stream: begin/end return elapsed ms via *ms. */
int tfcuda_timer_begin(void);
int tfcuda_timer_end(float* ms);

/* Per-kernel profiling.  While enabled every launch (emitted or library) is bracketed by a CUDA-event pair on the
 * runtime stream; records aggregate per kernel name.  `bytes` is the ALGORITHMIC traffic the caller attributes to the
 * launches (tfcuda_profile_add_bytes; the in-tree glue adds the size of every tensor bound to a dispatch, each once —
 * SURVEY.md §8d).  Used by bench.py for the roofline of the dominant kernel; off by default (no overhead). */
typedef struct TFCudaProfileRecord {
	char name[64];
	uint64_t launches;
	double total_ms;
	double bytes;
} TFCudaProfileRecord;
int tfcuda_profile_enable(int on);
int tfcuda_profile_reset(void);
int tfcuda_profile_add_bytes(size_t kernel_id, double bytes);
size_t tfcuda_profile_records(TFCudaProfileRecord* out, size_t capacity); /* syncs; returns the number of records */

/* Page-locked host memory for host<->device copies that run at full PCIe rate (bench.py's e2e leg). */
void* tfcuda_host_alloc(size_t bytes);
int tfcuda_host_free(void* p);

/* ------------------------------------------------------------------------------------------
 * Library kernels (hand-written sm_100a).  They replace the generic lowering of
 * Compiler/Implementations.cpp (ComputeReduction :243-303, ComputeScan :305-359, ComputeMatMul
 * :560-646), the user-level radix sort of Python/TensorFrost/sort.py:38-187, the CAS-loop float
 * atomics of CPP.cpp:150-158 and the fused all-pairs kernel of
 * examples/Simulation/n-body-benchmark.py:16-34.  All pointers are device addresses; all work is
 * enqueued on tfcuda_stream().
 * ------------------------------------------------------------------------------------------ */
enum {
	TFCUDA_RED_SUM = 0, TFCUDA_RED_MAX = 1, TFCUDA_RED_MIN = 2, TFCUDA_RED_MEAN = 3,
	TFCUDA_RED_NORM = 4, TFCUDA_RED_PROD = 5, TFCUDA_RED_ANY = 6, TFCUDA_RED_ALL = 7
};

/* out[o, i] = reduce_k in[o, k, i]   (in viewed as [outer, n, inner], row-major; the reference
 * reduces any axis: Implementations.cpp:243-303).  type is the element TFType. */
int tfcuda_reduce(uint64_t in, uint64_t out, size_t outer, size_t n, size_t inner, int op, int type);
/* inclusive prefix sum along the middle axis of [outer, n, inner] (Implementations.cpp:305-359) */
int tfcuda_prefix_sum(uint64_t in, uint64_t out, size_t outer, size_t n, size_t inner, int type);

/* Stable LSD radix sort of 32-bit keys (+ optional 32-bit values); key_type selects the key
 * bijection of sort.py:52-72 (TFUint: identity, TFInt: flip sign bit, TFFloat: IEEE total order).
 * keys_out/values_out receive the result; inputs are preserved.  values_* may be 0 (keys only).
 * temp: tfcuda_radix_sort_temp_words(n) words of scratch. */
size_t tfcuda_radix_sort_temp_words(size_t n);
int tfcuda_radix_sort(uint64_t keys_in, uint64_t keys_out, uint64_t values_in, uint64_t values_out,
                      size_t n, int key_type, int max_bits, uint64_t temp);

/* dst[index[i]] += src[i]  (tf.scatterAdd / autodiff of load: Implementations.cpp:185-194).
 * Warp-aggregated red.global.add; type TFFloat/TFInt/TFUint. */
int tfcuda_scatter_add(uint64_t dst, uint64_t index, uint64_t src, size_t n, size_t dst_words, int type);

/* C[b] = A[b] (MxK) @ B[b] (KxN), row-major fp32 (Implementations.cpp:560-646).
 * mode 0: tcgen05 kind::tf32 (1e-3 class); mode 1: 3xTF32 split (fp32-accurate); mode 2: FFMA. */
int tfcuda_matmul(uint64_t a, uint64_t b, uint64_t c, size_t batch, size_t m, size_t n, size_t k, int mode);

/* C (MxN) = A^T @ B with A [RxM], B [RxN] row-major fp32, contracted over the leading extent R: the weight gradient of
 * Y = X @ W (dW = X^T dY).  Replaces the reference VJP's batched Transpose(X)[b] @ dY[b] + batch reductions
 * (Implementations.cpp:133-135, 560-646, 243-303) with one split-K pass; deterministic (fixed-order partial sums). */
int tfcuda_matmul_tn(uint64_t a, uint64_t b, uint64_t c, size_t r, size_t m, size_t n);

/* EXPERIMENTAL (not yet run on hardware; off unless TFCUDA_MATMUL_ROWS=1): C (RxN) = A (RxK) @ B (KxN) for a small weight matrix
 * B (N <= 128, staged in shared memory once per persistent CTA) applied to very many rows; fp32 FFMA in the reference's k order
 * (Implementations.cpp:560-646), no TF32, no pre-pass over A.  tfcuda_matmul_rows_supported tells whether a shape qualifies. */
int tfcuda_matmul_rows_supported(size_t r, size_t k, size_t n);
int tfcuda_matmul_rows(uint64_t a, uint64_t b, uint64_t c, size_t r, size_t k, size_t n);

/* One all-pairs gravity step on N bodies, X,V: [N,3] fp32 (n-body-benchmark.py:16-34). */
int tfcuda_nbody_step(uint64_t x, uint64_t v, uint64_t x_new, uint64_t v_new, size_t n, float dt, float eps);

/* ------------------------------------------------------------------------------------------
 * Data-parallel exchange (new work; the reference has no distributed code):
 * one NCCL communicator per process, sum-allreduce of a flat fp32 buffer on tfcuda_stream().
 * unique_id: 128 bytes produced by rank 0 (tfcuda_comm_unique_id) and shared out of band.
 * ------------------------------------------------------------------------------------------ */
int tfcuda_comm_unique_id(uint8_t out[128]);
int tfcuda_comm_init(const uint8_t unique_id[128], int rank, int world);
int tfcuda_comm_allreduce_sum_f32(uint64_t buf, size_t count, float scale);
int tfcuda_comm_destroy(void);
/* One-shot allreduce over NVLink peer memory for small payloads (<= tfcuda_peer_max_count() floats; the NCA exchange is 7821):
 * every rank exports an exchange buffer (64-byte CUDA IPC handle, shared out of band like the NCCL id), maps its peers', and ONE
 * kernel per step pushes the vector into every peer's buffer, waits for all peers' flags and sums in rank order (bit-identical
 * result on every rank).  Same contract as tfcuda_comm_allreduce_sum_f32: in place on tfcuda_stream(), then multiplied by scale. */
int tfcuda_peer_export(uint8_t handle_out[64]);
int tfcuda_peer_init(const uint8_t* handles /* world x 64 bytes, rank order */, int rank, int world);
int tfcuda_peer_ready(void);
size_t tfcuda_peer_max_count(void);
int tfcuda_peer_allreduce_sum_f32(uint64_t buf, size_t count, float scale);
int tfcuda_peer_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* TFCUDA_H_ */
