// Glue between the reference's backend interfaces and libtfcuda.so (see CUDA.h).
// Errors are thrown as std::runtime_error, the reference's error contract (Backend/Backend.cpp:166-170).
#include <cstdlib>

#include "CUDA.h"

#define TFCUDA_NO_ABI_STRUCTS  // TFBuffer/TFTensor/... come from Backend/TensorMemory.h here
namespace tfcuda_abi {
using TensorFrost::TFBuffer;
using TensorFrost::TFDataFormat;
using TensorFrost::TFDispatchInfo;
using TensorFrost::TFRuntime;
using TensorFrost::TFTensor;
#include "tfcuda.h"
}  // namespace tfcuda_abi
using namespace tfcuda_abi;

namespace TensorFrost {

std::string cudaKernelCompileOptions;

static void Fail(const std::string& what) {
	throw std::runtime_error("CUDA backend: " + what + ": " + tfcuda_last_error());
}

void StartCUDA() {
	if (tfcuda_init(-1) != 0) Fail("initialisation failed (no CPU fallback)");
}

void StopCUDA() { tfcuda_sync(); }

void CudaFinish() {
	if (tfcuda_sync() != 0) Fail("stream synchronisation failed");
}

void CudaRegion(const char* name, bool begin) {
	TFRuntime rt = tfcuda_runtime();
	rt.region(name, begin, nullptr);
}

TFCudaBuffer::TFCudaBuffer(size_t size) : TFBufferTemplate(size) {
	TFBuffer* dev = tfcuda_buffer_create(size);
	if (dev == nullptr) Fail("cannot allocate " + std::to_string(size * 4) + " bytes");
	handle = dev;
	device_ptr = tfcuda_buffer_device_ptr(dev);
}

void TFCudaBuffer::Poison() {
	// TFCUDA_POISON=1: quiet-NaN pattern; TFCUDA_POISON=0x<word>: that word (e.g. 0x3f800000 = 1.0f, which unlike NaN passes `>` tests)
	static const char* env = getenv("TFCUDA_POISON");
	static const unsigned long word = env ? strtoul(env, nullptr, 0) : 0;
	if (word != 0 && tfcuda_memset32(device_ptr, word == 1 ? 0x7fc0dead : (uint32_t)word, size) != 0) Fail("poison fill failed");
}

TFCudaBuffer::~TFCudaBuffer() { tfcuda_buffer_destroy((TFBuffer*)handle); }

void TFCudaBuffer::SetDataAtOffset(size_t offset, const vector<uint32_t>& data) {
	if (offset + data.size() > size) throw std::runtime_error("CUDA backend: upload exceeds buffer " + std::string(name ? name : "?"));
	if (tfcuda_memcpy_h2d(device_ptr + offset * 4, data.data(), data.size() * 4) != 0) Fail("host to device copy failed");
}

void TFCudaBuffer::GetDataAtOffset(size_t offset, size_t count, uint32_t* data) {
	if (offset + count > size) throw std::runtime_error("CUDA backend: readback exceeds buffer " + std::string(name ? name : "?"));
	if (tfcuda_memcpy_d2h(data, device_ptr + offset * 4, count * 4) != 0) Fail("device to host copy failed");
}

void CudaKernelManager::CompileProgram(Program* program) {
	vector<TFCudaKernelSource> sources;
	sources.reserve(program->kernels_.size());
	for (auto& kernel : program->kernels_) {
		TFCudaKernelSource s{};
		s.kernel_id = kernel.kernel_id_;
		s.entry = kernel.kernel_name_.c_str();
		s.source = kernel.full_generated_code_.c_str();
		vector<int> group = kernel.root->group_size;
		while (group.size() < 3) group.push_back(1);
		for (int d = 0; d < 3; d++) s.group[d] = (unsigned)group[d];
		s.n_mem = (unsigned)kernel.GetMemoryBindings().size();
		s.n_var = (unsigned)kernel.var_names.size();
		s.library_op = FindCudaLibraryCall(kernel.kernel_id_) != nullptr ? 1 : 0;  // no source: dispatched by DispatchCudaLibraryCall
		sources.push_back(s);
	}
	if (tfcuda_compile_kernels(sources.data(), sources.size(), cudaKernelCompileOptions.c_str()) != 0) {
		Fail("kernel compilation failed for program " + program->program_name);
	}
}

void CudaKernelManager::DispatchKernel(TFDispatchInfo info) {
	if (const CudaLibraryCall* call = FindCudaLibraryCall(info.kernel_id)) {
		DispatchCudaLibraryCall(*call, info);
		return;
	}
	uint64_t ptrs[256];
	if (info.read_write_count > 256) throw std::runtime_error("CUDA backend: too many buffers in dispatch");
	for (size_t i = 0; i < info.read_write_count; i++) {
		ptrs[i] = ((TFCudaBuffer*)info.read_write_tensors[i].buffer)->GetNative();
	}
	// algorithmic traffic of this dispatch = every bound tensor once (only accounted while profiling)
	double bytes = 0;
	for (size_t i = 0; i < info.read_write_count; i++) bytes += 4.0 * (double)GetSize(&info.read_write_tensors[i]);
	tfcuda_profile_add_bytes(info.kernel_id, bytes);
	if (tfcuda_launch(info.kernel_id, ptrs, info.read_write_count, info.variables, info.variable_count, info.work_group_count) != 0) {
		Fail("dispatch of kernel " + std::to_string(info.kernel_id) + " failed");
	}
}

}  // namespace TensorFrost
