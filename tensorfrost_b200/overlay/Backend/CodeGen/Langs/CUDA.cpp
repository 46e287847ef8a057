// CUDA C++ emitter: lowers one fused IR kernel to an sm_100a __global__ function.
//
// Added to the reference tree as TensorFrost/Backend/CodeGen/Langs/CUDA.cpp; selected by
// `case CodeGenLang::CUDA` in GenerateKernel (Backend/CodeGen/Generators.cpp:9-23).  It plays the role
// GenerateGLSLKernel / GLSLGenerator play for OpenGL (Backend/CodeGen/Langs/GLSL.cpp:6-56,146-186) and
// GenerateCPPKernel plays for the OpenMP oracle (Backend/CodeGen/Langs/CPP.cpp:720-779): the per-node
// body text comes from the shared CodeGenerator::GenerateLine (Generators.cpp:319-540); this file owns
// the kernel signature, the binding/variable unpacking, shared memory, barriers and atomics.
//
// Emitted shape (contract with libtfcuda, include/tfcuda.h "Emitted kernel contract"):
//
//   struct kernel_<id>_args { uint* mem[n_mem]; uint var[n_var]; };       // n_var includes block offset
//   extern "C" __global__ void __launch_bounds__(G) kernel_<id>(const __grid_constant__ kernel_<id>_args tf_a)
//   {
//     __shared__ <type> <group_memory>[N]; ...
//     uint* <name>_mem = tf_a.mem[i]; ...            // rw bindings first, then ro (KernelGen.h:36-45)
//     <type> var_<name> = as<type>(tf_a.var[i]); ... // host scalars, declaration order of kernel->variables
//     int block_id = blockIdx.x + var__kernel_block_offset;      // 1-D grid (CPP.cpp:503-515)
//     TF_ASSUME(block_id >= 0);                                  // range fact for the compiler (prelude.cuh)
//     int block_thread_id{0,1,2} = threadIdx.{x,y,z};            // 0 = innermost
//     <body>
//   }
//
// The whole argument block travels in constant param space (one cuLaunchKernel parameter), so a
// dispatch needs no device-side pointer table and no UBO upload (cf. OpenGL/KernelManager.h:127-141).
#include <cstdlib>

#include "Backend/CodeGen/Generators.h"
#include "Backend/Backends/CUDA/CUDA.h"
#include "Backend/Backend.h"

namespace TensorFrost {
using namespace std;

namespace {

// IR function op -> prelude name.  Every function-class op gets a tf_ prefix so generated variable
// names (user debug names!) can never shadow a helper.
const char* const kFunctionOps[] = {
    "min",  "max",  "abs",  "sign", "ceil", "floor", "round", "frac",  "exp",   "exp2",        "log",  "log2",
    "sqrt", "rsqrt", "rcp", "sin",  "cos",  "tan",   "asin",  "acos",  "atan",  "sinh",        "cosh", "tanh",
    "pcg",  "pcgf", "pow",  "atan2", "modf", "step", "clamp", "lerp",  "fma",   "reversebits", "trunc", "smoothstep",
    "group_barrier"};

class CUDAGenerator : public CodeGenerator {
 public:
	explicit CUDAGenerator(IR* ir) : CodeGenerator(ir) {
		name_map_["var"] = "var_";
		for (const char* op : kFunctionOps) {
			name_map_[op] = string("tf_") + op;
		}
	}

	// Atomics address word buffers (global `<name>_mem` or a __shared__ array); the prelude overloads pick
	// the element type from the value, so the value is cast explicitly (Generators.h:183-189 casts the
	// pointer instead).
	string GenerateAtomicOp(const string& op, const string& input_type_name, const string& output_type_name,
	                        const string& address, const string& input, const string& output,
	                        const string& memory_name) override {
		static const unordered_map<string, string> fn = {
		    {"InterlockedAdd", "tf_atomic_add"}, {"InterlockedAdd_Prev", "tf_atomic_add_prev"},
		    {"InterlockedMin", "tf_atomic_min"}, {"InterlockedMax", "tf_atomic_max"},
		    {"InterlockedAnd", "tf_atomic_and"}, {"InterlockedOr", "tf_atomic_or"},
		    {"InterlockedXor", "tf_atomic_xor"}};
		auto it = fn.find(op);
		if (it == fn.end()) {
			throw runtime_error("CUDA emitter: unsupported atomic operation " + op);
		}
		return it->second + "((uint*)" + memory_name + ", " + address + ", (" + input_type_name + ")(" + input + "))";
	}
};

}  // namespace

// Default thread-block shape for kernels the user did not size (consulted by IR::LinearBlockModeIndices, Steps/GraphOps.cpp:1143-1166,
// which otherwise picks 256 / 16x16 / 8x8x8).  On a GPU the innermost kernel dimension is the contiguous one in memory, so a warp
// should span 32 consecutive innermost indices (one 128-byte line per row) instead of 16 or 8; the remaining threads go to the
// outer dimensions.  const_shape[i] > 0 where the extent is a compile-time constant (blocks never exceed it).  Returns {} for
// other kernel languages.
vector<int> CudaDefaultGroupSize(int dims, const vector<int>& const_shape) {
	if (current_kernel_lang != CodeGenLang::CUDA) return {};
	if (const char* v = getenv("TFCUDA_DEFAULT_GROUP")) {
		if (atoi(v) == 0) return {};  // debugging aid: the reference's own 256 / 16x16 / 8x8x8 defaults
	}
	auto pow2_floor = [](int v) { int p = 1; while (p * 2 <= v) p *= 2; return p; };
	auto extent = [&](int i) { return (i < (int)const_shape.size() && const_shape[i] > 0) ? const_shape[i] : (1 << 30); };
	const int target = 256;
	vector<int> group;
	int g0 = min(extent(0), dims == 1 ? target : 32);
	// A short constant innermost extent that is not a multiple of the warp (36 filter responses, 48 concatenated channels per cell in
	// the NCA example): take whole rows, so that consecutive thread ids are consecutive addresses across the rows of a block; 32-wide
	// blocks would leave a second block column with 4 (or 16) of 32 lanes busy.  TFCUDA_WHOLE_ROWS=0 restores the 32-wide choice.
	static const bool whole_rows = !(getenv("TFCUDA_WHOLE_ROWS") && atoi(getenv("TFCUDA_WHOLE_ROWS")) == 0);
	if (whole_rows && dims >= 2 && extent(0) > 32 && extent(0) <= 96 && extent(0) % 32 != 0) g0 = extent(0);
	group.push_back(g0);
	if (dims == 1) return group;
	int remaining = max(1, target / g0);
	if (dims == 2) {
		group.push_back(min(extent(1), pow2_floor(remaining)));
		return group;
	}
	int g1 = min(extent(1), max(1, pow2_floor(remaining) / 2));
	group.push_back(g1);
	remaining = max(1, target / (g0 * g1));
	group.push_back(min(min(extent(2), pow2_floor(remaining)), 64));
	return group;
}

namespace {

string CudaSharedDeclaration(const string& name, const string& type_name, int size) {
	return "  __shared__ " + type_name + " " + name + "[" + to_string(size) + "];\n";
}

}  // namespace

void GenerateCUDAKernel(Program* program, Kernel* kernel) {
	const string kname = kernel->kernel_name_;
	const string args_t = kname + "_args";

	// variable table: host scalars in declaration order + the block offset word (same layout as the
	// GLSL UBO, GLSL.cpp:96-113)
	kernel->var_names = vector<string>(kernel->variables.size());
	kernel->var_types = vector<string>(kernel->variables.size());
	for (auto var : kernel->variables) {
		kernel->var_names[var.second] = var.first->var_name;
		kernel->var_types[var.second] = type_names[var.first->format.type];
	}
	kernel->var_names.push_back("_kernel_block_offset");
	kernel->var_types.push_back(type_names[TFType::Uint]);

	const size_t n_mem = kernel->GetMemoryBindings().size();
	const size_t n_var = kernel->var_names.size();

	vector<int> group = kernel->root->group_size;
	while (group.size() < 3) group.push_back(1);
	const int threads = group[0] * group[1] * group[2];

	if (IsCudaLibraryKernel(kernel)) {
		// a library call (CudaLibrary.cpp): record which binding plays which role; there is no source to emit
		RegisterCudaLibraryKernel(kernel);
		kernel->generated_header_ = "";
		kernel->generated_bindings_ = "";
		// the comment also records which binding plays which role (tests/cpu_sim replays library-lowered programs on the host from it)
		string roles;
		if (const CudaLibraryCall* call = FindCudaLibraryCall(kernel->kernel_id_)) {
			roles = " inputs=[";
			for (size_t i = 0; i < call->inputs.size(); i++) roles += (i ? "," : "") + to_string(call->inputs[i]);
			roles += "] outputs=[";
			for (size_t i = 0; i < call->outputs.size(); i++) roles += (i ? "," : "") + to_string(call->outputs[i]);
			roles += "]";
		}
		kernel->generated_main_ = "// " + kname + ": " + kernel->root->debug_name + roles + " -> hand-written kernel in libtfcuda.so\n";
		kernel->full_generated_code_ = kernel->generated_main_;
		return;
	}

	// the body first: generating it may rename nodes (RegenerateNodeName), and binding names below must
	// be read after that
	CUDAGenerator generator(program->ir_);
	generator.GenerateKernelCode(kernel);
	const string body = generator.AssembleString();

	string bindings = "struct " + args_t + " {\n";
	if (n_mem > 0) bindings += "  uint* mem[" + to_string(n_mem) + "];\n";
	bindings += "  uint var[" + to_string(n_var) + "];\n};\n";

	// TFCUDA_MIN_BLOCKS=n (tuning aid): ask for n resident blocks per SM, i.e. cap the registers per thread at 65536 / (n * threads)
	string bounds = to_string(threads);
	if (const char* mb = getenv("TFCUDA_MIN_BLOCKS")) {
		if (atoi(mb) > 0) bounds += ", " + to_string(atoi(mb));
	}
	string main_code = "extern \"C\" __global__ void __launch_bounds__(" + bounds + ") " + kname +
	                   "(const __grid_constant__ " + args_t + " tf_a)\n{\n";
	main_code += GetGroupBufferDeclarations(kernel, CudaSharedDeclaration);
	// read-only bindings (Kernel::read_only_memory come after the rw ones, KernelGen.h:36-45) are declared through TF_RO
	// (= const uint* __restrict__, prelude.cuh): loads take the non-coherent path and may be hoisted / batched by the compiler
	const char* ro_env = getenv("TFCUDA_RO");  // debugging aid: TFCUDA_RO=0 declares every binding as plain uint*
	const size_t n_rw = (ro_env && atoi(ro_env) == 0) ? n_mem : kernel->read_write_memory.size();
	main_code += GetBufferDeclarations(kernel, [n_rw](const string& name, const string& type_name, size_t binding) {
		const string type = binding >= n_rw ? "TF_RO " : "uint* ";
		return "  " + type + name + "_mem = tf_a.mem[" + to_string(binding) + "];\n";
	});
	for (size_t i = 0; i < n_var; i++) {
		main_code += "  " + kernel->var_types[i] + " var_" + kernel->var_names[i] + " = as" + kernel->var_types[i] +
		             "(tf_a.var[" + to_string(i) + "]);\n";
	}
	if (const char* pdl = getenv("TFCUDA_PDL")) {
		if (atoi(pdl) != 0) main_code += "  tf_pdl_prologue();\n";  // experimental: programmatic dependent launch (prelude.cuh)
	}
	main_code += "  int block_id = (int)(blockIdx.x + var__kernel_block_offset);\n";
	// a fact the compiler cannot derive (prelude.cuh): lets it drop the `>= 0` half of index clamps and bounds tests.
	// TFCUDA_ASSUME=0 (debugging aid) leaves it out.
	static const bool assume_env = !(getenv("TFCUDA_ASSUME") && atoi(getenv("TFCUDA_ASSUME")) == 0);
	if (assume_env) main_code += "  TF_ASSUME(block_id >= 0);\n";
	main_code += "  int block_thread_id0 = (int)threadIdx.x;\n";
	main_code += "  int block_thread_id1 = (int)threadIdx.y;\n";
	main_code += "  int block_thread_id2 = (int)threadIdx.z;\n";
	main_code += "  (void)block_id; (void)block_thread_id0; (void)block_thread_id1; (void)block_thread_id2;\n\n";
	main_code += AddIndent(body, "  ");
	main_code += "}\n";

	kernel->generated_header_ = "";  // the prelude is shared by all kernels: libtfcuda prepends it per NVRTC unit
	kernel->generated_bindings_ = bindings;
	kernel->generated_main_ = main_code;
	kernel->full_generated_code_ = bindings + main_code;
}

}  // namespace TensorFrost
