// CUDA (B200, sm_100a) backend for TensorFrost: the third sibling of Backends/CPU and Backends/OpenGL.
//
// Added to the reference tree as TensorFrost/Backend/Backends/CUDA/CUDA.h.  It is deliberately thin: all
// device work lives behind the C-ABI of libtfcuda.so (include/tfcuda.h); this header only adapts that ABI
// to the reference's three C++ extension points:
//   TensorMemoryManager::{CreateBuffer,DeleteBuffer}          Backend/TensorMemory.h:112-125
//   TFBufferTemplate::{UpdateName,SetDataAtOffset,GetDataAtOffset}   Backend/TensorMemory.h:74-88
//   KernelManager::DispatchKernel (+ a CompileProgram hook)   Backend/KernelManager.h:16-30
// mirroring what Backends/CPU/{Memory.h,KernelManager.h} and Backends/OpenGL/{Memory.h,KernelManager.h} do.
#pragma once

#include <array>
#include <string>
#include <vector>

#include "../../KernelManager.h"
#include "../../TensorMemory.h"
#include "Tensor/Tensor.h"

namespace TensorFrost {

using namespace std;

// NVRTC flags for emitted kernels (the kernel_compile_options string of tf.initialize)
extern std::string cudaKernelCompileOptions;
// value of a --tf-<name>=<value> token in that string (backend options, stripped before NVRTC sees the rest)
std::string CudaBackendOption(const std::string& name, const std::string& fallback);

void StartCUDA();
void StopCUDA();
void CudaFinish();
void CudaRegion(const char* name, bool begin);
// Brackets one execution of a compiled program (ExecuteProgram, Backend/Backend.cpp:137-185): between Begin and End the runtime
// records the program's dispatches and replays them as one CUDA graph (tfcuda_graph_begin / tfcuda_graph_end).  No-op on other backends.
void CudaProgramBegin();
void CudaProgramEnd(bool may_throw);
struct CudaProgramScope {
	bool open = true;
	CudaProgramScope() { CudaProgramBegin(); }
	void Finish() { open = false; CudaProgramEnd(true); }
	~CudaProgramScope() { if (open) CudaProgramEnd(false); }
};

// first statement of CompileKernelLibrary (Backends/CPU/KernelCompiler.cpp:93-113) in the overlay build: when it returns true the
// compiled host program is already at `dll_name` (per-process file names + a content-addressed cache instead of the fixed
// /tmp/generated_lib_<id>.cpp); false = not our backend, the reference path runs
bool CudaHostProgramCache(const std::string& code, const char* dll_name, size_t program_id);

// ---- library calls inside compiled programs (CudaLibrary.cpp) ---------------------------------------------------------
// One hand-written libtfcuda kernel standing in for a lowered algorithmic op; `inputs` / `outputs` are binding indices into
// the dispatch's tensor list (rw bindings first, then ro: Compiler/KernelGen.h:36-45).
struct CudaLibraryCall {
	std::string op;            // "reduce" | "scan" | "matmul" | "sort"
	std::vector<int> params;   // reduce: {TFCUDA_RED_*, internal axis}; scan: {internal axis}; matmul: {mode}; sort: {has_values, max_bits}
	std::vector<int> inputs;
	std::vector<int> outputs;
};
void InstallCudaLibraryLowerings();                     // replaces entries of implementation_functions (Compiler/Implementations.cpp:648)
bool CudaLibraryWantsReduction(Node* node);             // consulted by IR::OptimizeReductions (Steps/Optimization.cpp:471) before it stages a reduction
std::vector<Tensor*> CudaLibrarySort(const Tensor* keys, const Tensor* values, int max_bits);
bool IsCudaLibraryKernel(Kernel* kernel);
// threads per block an emitted kernel is launched with: the IR's group_size, except for coarsened kernels (CodeGen/Langs/CUDA.cpp)
std::array<int, 3> CudaLaunchBlock(const Kernel* kernel);
void RegisterCudaLibraryKernel(Kernel* kernel);
const CudaLibraryCall* FindCudaLibraryCall(size_t kernel_id);
void DispatchCudaLibraryCall(const CudaLibraryCall& call, const TFDispatchInfo& info);

class TFCudaBuffer : public TFBufferTemplate {
 public:
	uint64_t device_ptr = 0;
	void* handle = nullptr;  // TFBuffer* owned by libtfcuda

	explicit TFCudaBuffer(size_t size);
	~TFCudaBuffer();

	// called by TensorMemoryManager::AllocateTensor for every (re)allocation (Backend/TensorMemory.cpp:43-56): also the hook of
	// the TFCUDA_POISON debugging aid, which fills every newly handed-out buffer with a NaN pattern so that programs reading
	// memory they never wrote fail deterministically instead of depending on what the pool recycled
	void UpdateName(const char* new_name) override {
		if (new_name != nullptr) name = new_name;
		Poison();
	}
	void Poison();
	void SetDataAtOffset(size_t offset, const vector<uint32_t>& data) override;
	void GetDataAtOffset(size_t offset, size_t size, uint32_t* data) override;
	uint64_t GetNative() const { return device_ptr; }
};

class CudaMemoryManager : public TensorMemoryManager {
 public:
	TFBuffer* CreateBuffer(size_t size) override { return new TFCudaBuffer(size); }
	void DeleteBuffer(TFBuffer* buffer) override { delete (TFCudaBuffer*)buffer; }
};

class CudaKernelManager : public KernelManager {
 public:
	// compiles every kernel of the program in one call (parallel NVRTC chunks inside libtfcuda)
	void CompileProgram(Program* program);
	void DispatchKernel(TFDispatchInfo info) override;
};

}  // namespace TensorFrost
