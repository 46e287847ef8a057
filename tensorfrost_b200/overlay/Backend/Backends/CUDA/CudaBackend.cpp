// Glue between the reference's backend interfaces and libtfcuda.so (see CUDA.h).
// Errors are thrown as std::runtime_error, the reference's error contract (Backend/Backend.cpp:166-170).
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <unordered_map>

#include "CUDA.h"
#include "Backend/Backend.h"
#include "Backend/Backends/CPU/KernelCompiler.h"

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

// Backend options travel in the kernel_compile_options string of tf.initialize (Frontend/Python/PybindModule.cpp:112-115) next to
// the NVRTC flags, as tokens of the form --tf-<name>=<value>; e.g. tf.initialize(tf.cuda, "--tf-matmul=tf32").
//   --tf-matmul=3xtf32 | tf32 | fp32    precision of `a @ b` in compiled programs (default 3xtf32: fp32-accurate; tf32: one tensor-core
//                                       product, 1e-3 class, ~3x faster; fp32: FFMA kernel)
std::string CudaBackendOption(const std::string& name, const std::string& fallback) {
	const std::string key = "--tf-" + name + "=";
	size_t at = cudaKernelCompileOptions.find(key);
	if (at == std::string::npos) return fallback;
	size_t end = cudaKernelCompileOptions.find_first_of(" \t", at);
	return cudaKernelCompileOptions.substr(at + key.size(), end == std::string::npos ? std::string::npos : end - at - key.size());
}

static std::string NvrtcOptions() {
	std::string out;
	size_t i = 0;
	const std::string& s = cudaKernelCompileOptions;
	while (i < s.size()) {
		size_t j = s.find_first_of(" \t", i);
		if (j == std::string::npos) j = s.size();
		std::string tok = s.substr(i, j - i);
		if (!tok.empty() && tok.rfind("--tf-", 0) != 0) out += (out.empty() ? "" : " ") + tok;
		i = j + 1;
	}
	return out;
}

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

void CudaProgramBegin() {
	if (current_backend == BackendType::CUDA) tfcuda_graph_begin();
}

void CudaProgramEnd(bool may_throw) {
	if (current_backend != BackendType::CUDA) return;
	if (tfcuda_graph_end() != 0 && may_throw) Fail("a kernel launch of the program failed");
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

// ---- host program: compile once, cache by content, never share a file name between processes -------------------------
// The reference writes every program's host code to the FIXED path /tmp/generated_lib_<program id>.cpp and compiles it with a
// g++ subprocess on every tf.compile (Backends/CPU/KernelCompiler.cpp:93-113).  Two ranks of a multi-GPU job tracing the same
// program therefore overwrite each other's source while g++ reads it, and every process pays 0.5-7 s per program again.
// Here the library is keyed by the hash of the generated text + compiler flags: a hit costs a symlink; a miss compiles from
// <hash>.<pid>.<id>.cpp to <hash>.<pid>.<id>.so and renames it into place, so concurrent ranks never see a partial file.
// `dll_name` is the unique path the reference chose with mktemp and dlopens next (KernelCompiler.cpp:135-153).
// Active on the CUDA backend (and, for the CPU tests of this very function, on any backend when TFCUDA_HOST_CACHE=1).
namespace {

// single-quote a path for the shell command below (the cache directory comes from the environment)
std::string ShellQuote(const std::string& s) {
	std::string out = "'";
	for (char c : s) {
		if (c == '\'') out += "'\\''";
		else out += c;
	}
	return out + "'";
}

uint64_t Fnv1a(const std::string& s, uint64_t h) {
	for (unsigned char c : s) {
		h ^= c;
		h *= 1099511628211ull;
	}
	return h;
}

}  // namespace

bool CudaHostProgramCache(const std::string& code, const char* dll_name, size_t program_id) {
	const char* force = getenv("TFCUDA_HOST_CACHE");
	if (current_backend != BackendType::CUDA && !(force && atoi(force) != 0)) return false;
	const std::string flags = kernelCompileOptions;
	const std::string key = code + "\x01" + flags;
	char name[80];
	snprintf(name, sizeof(name), "host_%016llx%016llx", (unsigned long long)Fnv1a(key, 1469598103934665603ull),
	         (unsigned long long)Fnv1a(key, 0x9e3779b97f4a7c15ull));
	const std::string cache = tfcuda_cache_dir();
	const bool cached = !cache.empty() && !(force && atoi(force) == 0);
	const std::string final_so = cached ? cache + "/" + name + ".so" : std::string(dll_name);
	struct stat st;
	if (!(cached && stat(final_so.c_str(), &st) == 0 && st.st_size > 0)) {
		const std::string unique = (cached ? cache + "/" + name : std::string(dll_name)) + "." + std::to_string((long)getpid()) + "." + std::to_string(program_id);
		const std::string src = unique + ".cpp", tmp_so = unique + ".so";
		{
			std::ofstream out(src);
			if (!out) throw std::runtime_error("CUDA backend: cannot write the generated host program to " + src);
			out << code;
		}
		const std::string cmd = "g++ " + flags + " -w -shared -fPIC " + ShellQuote(src) + " -o " + ShellQuote(tmp_so) + " 2>&1";
		std::string output;
		FILE* pipe = popen(cmd.c_str(), "r");
		if (!pipe) throw std::runtime_error("CUDA backend: popen(g++) failed");
		char buffer[256];
		while (fgets(buffer, sizeof(buffer), pipe) != nullptr) output += buffer;
		int status = pclose(pipe);
		if (!getenv("TFCUDA_KEEP_HOST_SOURCE")) unlink(src.c_str());
		if (status != 0) {
			unlink(tmp_so.c_str());
			throw std::runtime_error("CUDA backend: host program compiler exited with status " + std::to_string(status) + "\nCompiler output:\n" + output);
		}
		if (rename(tmp_so.c_str(), final_so.c_str()) != 0) {
			unlink(tmp_so.c_str());
			throw std::runtime_error("CUDA backend: cannot move the compiled host program into " + final_so);
		}
	}
	if (cached && symlink(final_so.c_str(), dll_name) != 0) {
		throw std::runtime_error(std::string("CUDA backend: cannot link the cached host program to ") + dll_name);
	}
	return true;
}

// kernel id -> number of read-write bindings (the read-only ones follow, Compiler/KernelGen.h:36-45).  Read-only bindings are declared
// `const uint* __restrict__` in emitted kernels (a no-alias promise to the compiler); DispatchKernel checks the promise.
static std::unordered_map<size_t, size_t> g_rw_count;

void CudaKernelManager::CompileProgram(Program* program) {
	for (auto& kernel : program->kernels_) g_rw_count[kernel.kernel_id_] = kernel.read_write_memory.size();
	vector<TFCudaKernelSource> sources;
	sources.reserve(program->kernels_.size());
	for (auto& kernel : program->kernels_) {
		TFCudaKernelSource s{};
		s.kernel_id = kernel.kernel_id_;
		s.entry = kernel.kernel_name_.c_str();
		s.source = kernel.full_generated_code_.c_str();
		const std::array<int, 3> block = CudaLaunchBlock(&kernel);  // != kernel.root->group_size for coarsened kernels
		for (int d = 0; d < 3; d++) s.group[d] = (unsigned)block[d];
		s.n_mem = (unsigned)kernel.GetMemoryBindings().size();
		s.n_var = (unsigned)kernel.var_names.size();
		s.library_op = FindCudaLibraryCall(kernel.kernel_id_) != nullptr ? 1 : 0;  // no source: dispatched by DispatchCudaLibraryCall
		sources.push_back(s);
	}
	if (tfcuda_compile_kernels(sources.data(), sources.size(), NvrtcOptions().c_str()) != 0) {
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
	// A read-only binding that is the same device buffer as a writable one (a reshape view of a tensor the kernel also writes, or one
	// TensorMemory passed as two program inputs) would break the __restrict__ promise of the emitted kernel: refuse loudly.
	auto rw = g_rw_count.find(info.kernel_id);
	static const bool no_restrict = cudaKernelCompileOptions.find("TF_NO_RESTRICT") != std::string::npos || getenv("TFCUDA_RO") != nullptr;
	if (rw != g_rw_count.end() && !no_restrict) {
		for (size_t i = rw->second; i < info.read_write_count; i++)
			for (size_t j = 0; j < rw->second && j < info.read_write_count; j++)
				if (ptrs[i] == ptrs[j])
					throw std::runtime_error("CUDA backend: kernel " + std::to_string(info.kernel_id) + " binds one device buffer both read-only and writable; "
					                         "initialise with tf.initialize(tf.cuda, \"-DTF_NO_RESTRICT\") (or TFCUDA_RO=0 at trace time) for such programs");
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
