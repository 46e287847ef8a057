// Internal shared state of libtfcuda.so (not part of the C-ABI).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../../include/tfcuda.h"

namespace tfcuda {

// Driver-API entry points are resolved through cudaGetDriverEntryPoint so the library has no link-time
// dependency on libcuda.so.1 (it must load — and export its symbols — on a box without a driver).
struct DriverApi {
	CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
	CUresult (*ModuleUnload)(CUmodule) = nullptr;
	CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
	CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned,
	                         unsigned, CUstream, void**, void**) = nullptr;
	CUresult (*LaunchKernelEx)(const CUlaunchConfig*, CUfunction, void**, void**) = nullptr;  // optional (TFCUDA_PDL)
	CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
	CUresult (*FuncGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
	// graph replay of recorded launches (runtime.cu "Launch recorder"); optional: without them launches stay eager
	CUresult (*GraphCreate)(CUgraph*, unsigned) = nullptr;
	CUresult (*GraphAddKernelNode)(CUgraphNode*, CUgraph, const CUgraphNode*, size_t, const CUDA_KERNEL_NODE_PARAMS*) = nullptr;
	CUresult (*GraphInstantiate)(CUgraphExec*, CUgraph, unsigned long long) = nullptr;
	CUresult (*GraphLaunch)(CUgraphExec, CUstream) = nullptr;
	CUresult (*GraphExecKernelNodeSetParams)(CUgraphExec, CUgraphNode, const CUDA_KERNEL_NODE_PARAMS*) = nullptr;
	CUresult (*GraphExecDestroy)(CUgraphExec) = nullptr;
	CUresult (*GraphDestroy)(CUgraph) = nullptr;
	CUresult (*GraphAddDependencies_v2)(CUgraph, const CUgraphNode*, const CUgraphNode*, const CUgraphEdgeData*, size_t) = nullptr;
};

struct State {
	bool initialized = false;
	int device = -1;
	int sm_count = 0;
	std::string device_name;
	cudaStream_t stream = nullptr;
	DriverApi drv;
	uint32_t* pinned_word = nullptr;  // 1-word staging for tf.read / tf.write
	cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
	uint64_t launches = 0;
};

State& state();
void set_error(const std::string& msg);
std::string cuda_err(cudaError_t e);

// The device buffer behind a TFBuffer: TFBuffer must stay the first member so TFBuffer* <-> Buffer* is a cast.
struct Buffer {
	TFBuffer base;
	uint64_t dptr = 0;
};

inline uint64_t dptr_of(const TFBuffer* b) { return reinterpret_cast<const Buffer*>(b)->dptr; }

void require_init();  // throws std::runtime_error when tfcuda_init has not succeeded

// Grow-only device scratch shared by the library kernels (matmul planes, split-K partials).  Everything runs on ONE stream, so
// consecutive library calls can reuse the same block; keeping it avoids a multi-GB cudaMallocAsync/cudaFreeAsync pair per call
// (150 per NCA training step).  Returns nullptr (error set) when the device cannot provide it.  Valid until the next call.
void* scratch(size_t bytes);

#define TFCUDA_CHECK(expr)                                                                   \
	do {                                                                                     \
		cudaError_t _e = (expr);                                                             \
		if (_e != cudaSuccess) {                                                             \
			::tfcuda::set_error(std::string(#expr) + ": " + ::tfcuda::cuda_err(_e));         \
			return 1;                                                                        \
		}                                                                                    \
	} while (0)

#define TFCUDA_THROW(expr)                                                                   \
	do {                                                                                     \
		cudaError_t _e = (expr);                                                             \
		if (_e != cudaSuccess) {                                                             \
			std::string _m = std::string("tfcuda: ") + #expr + ": " + ::tfcuda::cuda_err(_e); \
			::tfcuda::set_error(_m);                                                         \
			throw std::runtime_error(_m);                                                    \
		}                                                                                    \
	} while (0)

// Profiling hooks (runtime.cu): a ProfileScope brackets the launches of one named kernel with an event pair when
// profiling is enabled, and is free otherwise.
void profile_begin(const char* name);
void profile_end(const char* name, double bytes);
struct ProfileScope {
	const char* name;
	double bytes;
	ProfileScope(const char* n, double b = 0.0) : name(n), bytes(b) { profile_begin(name); }
	~ProfileScope() { profile_end(name, bytes); }
};

// library kernels report launch errors through this
inline int check_launch(const char* what) {
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) {
		set_error(std::string(what) + ": " + cuda_err(e));
		return 1;
	}
	state().launches++;
	return 0;
}

}  // namespace tfcuda
