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
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unordered_set>

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

// ---- thread coarsening (SURVEY.md 2.2 / VERDICT r1 "emitter that moves more than 4 B per thread") ---------------------------------
// A streaming kernel with one element per thread keeps one 4-byte load per input in flight per thread: ~8 KB per SM at full occupancy where
// Little's law asks for ~35 KB at HBM3e rate and latency (measured round 2: NCA's bias + LeakyReLU kernel 2.6 TB/s, fluid's Jacobi and
// gradient kernels 41-56 % of HBM).  For a kernel that (i) got the DEFAULT block shape, (ii) is large and fully constant in shape and
// (iii) has a straight-line body, the block the IR sees (`group_size`, which the index arithmetic and the host's block count are built
// from) is made `factor` times larger along one dimension, the kernel is LAUNCHED with the original block, and every CUDA thread carries
// `factor` "lanes": the emitter replicates the body statement by statement (lane 0's copy of statement 1, lane 1's copy of statement 1,
// ..., then statement 2), so all lanes' loads are issued before the first dependent arithmetic and the stores come last.  Everything a
// lane shares with its neighbours (row offsets, uniform loads, for adjacent lanes the common stencil taps) is one value to the compiler.
struct CudaCoarsening {
	int dim = 0;      // group dimension the lanes run along
	int factor = 1;   // lanes per CUDA thread
	bool exact = false;  // every constant extent is a multiple of the virtual block: the dispatch guard is always true
	vector<int> group;   // the enlarged block handed to the IR (checked again at emission: the table is keyed by node address)
};
static unordered_map<const Node*, CudaCoarsening> g_coarsened;   // kernel node -> decision taken when its default block was chosen
static unordered_map<size_t, array<int, 3>> g_launch_block;      // kernel id -> threads per block the kernel is launched with

static int CudaCoarsenFactor() {
	static const int factor = [] {
		const char* v = getenv("TFCUDA_COARSEN");  // lanes per thread; 0 or 1 switches coarsening off
		int f = v ? atoi(v) : 4;
		return (f == 2 || f == 4 || f == 8) ? f : 1;
	}();
	return factor;
}

// straight-line body (loops every lane runs in step are allowed: the emitter checks their headers), few memory operations, nothing that
// ties the block shape to the program (barriers, group memory, thread ids)
static bool CudaKernelIsCoarsenable(Node* kernel_node) {
	int nodes = 0, memory_ops = 0;
	for (auto node = NodeIterator(kernel_node); !node.end(); node.next()) {
		const Operation* op = node->op;
		if ((op->HasAllTypes(OpProp::HasChildren) && op->name_ != "loop") || op->class_ == OpClass::Keyword || op->HasAllTypes(OpProp::LocalMemory) ||
		    node->flags.has(NodeProp::LocalMemoryOp) || op->name_ == "group_barrier" || op->name_ == "block_thread_id") {
			if (getenv("TFCUDA_COARSEN_DEBUG")) fprintf(stderr, "[tfcuda coarsen] refused: %s\n", op->name_.c_str());
			return false;
		}
		if (op->HasAllTypes(OpProp::MemoryOp)) memory_ops++;
		nodes++;
	}
	static const bool debug = getenv("TFCUDA_COARSEN_DEBUG") != nullptr;
	if (debug) fprintf(stderr, "[tfcuda coarsen] %s: %d nodes, %d memory ops\n", kernel_node->debug_name.c_str(), nodes, memory_ops);
	static const int max_memory_ops = getenv("TFCUDA_COARSEN_MAX_MEMOPS") ? atoi(getenv("TFCUDA_COARSEN_MAX_MEMOPS")) : 16;  // tuning aid; measured: the fluid projection
	// kernel (28 memory operations, 64 registers with 4 lanes) gained nothing from lanes (64.9 -> 65.7 us)
	static const int max_nodes = getenv("TFCUDA_COARSEN_MAX_NODES") ? atoi(getenv("TFCUDA_COARSEN_MAX_NODES")) : 240;
	return nodes > 0 && nodes <= max_nodes && memory_ops <= max_memory_ops;
}

// Default thread-block shape for kernels the user did not size (consulted by IR::LinearBlockModeIndices, Steps/GraphOps.cpp:1143-1166,
// which otherwise picks 256 / 16x16 / 8x8x8).  On a GPU the innermost kernel dimension is the contiguous one in memory, so a warp
// should span 32 consecutive innermost indices (one 128-byte line per row) instead of 16 or 8; the remaining threads go to the
// outer dimensions.  const_shape[i] > 0 where the extent is a compile-time constant (blocks never exceed it).  Returns {} when
// the reference's own defaults are asked for.
static vector<int> CudaBaseGroupSize(int dims, const vector<int>& const_shape) {
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

// The block IR::LinearBlockModeIndices uses for a kernel the user did not size: the shape above, enlarged along one dimension when the
// kernel is going to be coarsened.  Returns {} for other kernel languages (the reference's defaults apply).
vector<int> CudaDefaultGroupSize(int dims, const vector<int>& const_shape, Node* kernel_node) {
	if (current_kernel_lang != CodeGenLang::CUDA) return {};
	if (kernel_node != nullptr) g_coarsened.erase(kernel_node);  // a recycled node address must not inherit an old decision
	vector<int> group = CudaBaseGroupSize(dims, const_shape);
	const int factor = CudaCoarsenFactor();
	if (group.empty() || factor <= 1 || kernel_node == nullptr) return group;
	// large, constant shape only: the grid must still fill the machine after losing `factor` of its blocks (148 SMs x 8 blocks of 256
	// threads = 1184 blocks per wave), and a short kernel is launch-bound, not bandwidth-bound.  The bar is 2^20 elements: 1024 blocks
	// of 256 threads, one full wave (measured at 2^22: the fluid stencil kernels -25 .. -41 %; the 1024^2 multigrid level is the same code
	// at 2^20, where the issue time of one wave halves; at 2^18 the grid would leave SMs idle)
	long long elements = 1;
	for (int i = 0; i < dims; i++) {
		if (i >= (int)const_shape.size() || const_shape[i] <= 0) return group;
		elements *= const_shape[i];
	}
	// TFCUDA_COARSEN_MIN_ELEMENTS: the threshold, so that the host-execution tests can run their small cases through the lane code
	static const long long min_elements = getenv("TFCUDA_COARSEN_MIN_ELEMENTS") ? atoll(getenv("TFCUDA_COARSEN_MIN_ELEMENTS")) : (1ll << 20);
	if (elements < min_elements) return group;
	int threads = 1;
	for (int g : group) threads *= g;
	if (threads * factor > 1024) return group;  // the virtual block must stay launchable as it is (fallback of the emitter)
	if (!CudaKernelIsCoarsenable(kernel_node)) return group;
	// lanes along the first outer dimension with room for them (rows: adjacent lanes share stencil taps and stay coalesced); a 1-D kernel
	// takes lanes a whole block apart so that every load instruction still covers consecutive addresses
	int dim = -1;
	for (int d = 1; d < (int)group.size() && dim < 0; d++) {
		if (const_shape[d] >= group[d] * factor) dim = d;
	}
	if (dim < 0 && const_shape[0] >= group[0] * factor && (dims == 1 || group[0] % 32 == 0)) dim = 0;
	if (dim < 0) return group;
	group[dim] *= factor;
	CudaCoarsening c;
	c.dim = dim;
	c.factor = factor;
	c.exact = true;
	for (int d = 0; d < (int)group.size(); d++) c.exact = c.exact && (const_shape[d] % group[d] == 0);
	c.group = group;
	g_coarsened[kernel_node] = c;
	return group;
}

array<int, 3> CudaLaunchBlock(const Kernel* kernel) {
	auto it = g_launch_block.find(kernel->kernel_id_);
	if (it != g_launch_block.end()) return it->second;
	vector<int> group = kernel->root->group_size;
	while (group.size() < 3) group.push_back(1);
	return {group[0], group[1], group[2]};
}

namespace {

// ---- lane replication of an emitted body (see "thread coarsening" above) -------------------------------------------------------
// The body the shared generator produced has the shape
//     <declarations: constants, block decomposition, index_k>          "prologue"
//     bool is_inside_dispatch = ...;
//     if (is_inside_dispatch)
//     {
//       <straight-line statements>                                     "payload"
//     }
// Returns "" when the text is not of that shape (the caller then falls back to a lane loop around the untouched body).
vector<string> SplitLines(const string& text) {
	vector<string> lines;
	size_t at = 0;
	while (at < text.size()) {
		size_t end = text.find('\n', at);
		if (end == string::npos) end = text.size();
		lines.push_back(text.substr(at, end - at));
		at = end + 1;
	}
	return lines;
}

string Trimmed(const string& line) {
	size_t a = line.find_first_not_of(" \t");
	if (a == string::npos) return "";
	size_t b = line.find_last_not_of(" \t");
	return line.substr(a, b - a + 1);
}

bool IsIdentStart(char c) { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; }
bool IsIdentChar(char c) { return IsIdentStart(c) || (c >= '0' && c <= '9'); }

// a plain statement: no braces, no control flow, ends with ';'
bool IsPlainStatement(const string& t) {
	if (t.empty() || t.back() != ';') return false;
	if (t.find('{') != string::npos || t.find('}') != string::npos) return false;
	static const char* const kKeywords[] = {"if", "for", "while", "else", "do", "switch", "break", "continue", "return", "discard", "goto", "case"};
	size_t n = 0;
	while (n < t.size() && IsIdentChar(t[n])) n++;
	const string first = t.substr(0, n);
	for (const char* k : kKeywords) if (first == k) return false;
	return true;
}

// `<type> <name> = ...;` -> name
string DeclaredName(const string& t) {
	static const char* const kTypes[] = {"float ", "int ", "uint ", "bool "};
	for (const char* type : kTypes) {
		const size_t n = strlen(type);
		if (t.compare(0, n, type) != 0) continue;
		size_t a = n;
		while (a < t.size() && t[a] == ' ') a++;
		size_t b = a;
		while (b < t.size() && IsIdentChar(t[b])) b++;
		if (b == a || !IsIdentStart(t[a])) return "";
		size_t c = b;
		while (c < t.size() && t[c] == ' ') c++;
		if (c < t.size() && t[c] == '=' && (c + 1 >= t.size() || t[c + 1] != '=')) return t.substr(a, b - a);
		return "";
	}
	return "";
}

// every identifier of `names` gets `suffix`; numeric literals (1.f, 0x7fu, 1e-5f) are skipped as a whole
string RenameIdentifiers(const string& line, const unordered_set<string>& names, const string& suffix) {
	string out;
	out.reserve(line.size() + 32);
	size_t i = 0;
	while (i < line.size()) {
		const char c = line[i];
		if (IsIdentStart(c)) {
			size_t j = i;
			while (j < line.size() && IsIdentChar(line[j])) j++;
			const string token = line.substr(i, j - i);
			out += token;
			if (names.count(token)) out += suffix;
			i = j;
		} else if ((c >= '0' && c <= '9') || (c == '.' && i + 1 < line.size() && line[i + 1] >= '0' && line[i + 1] <= '9')) {
			size_t j = i;
			while (j < line.size() && (IsIdentChar(line[j]) || line[j] == '.')) j++;
			out += line.substr(i, j - i);
			i = j;
		} else {
			out += c;
			i++;
		}
	}
	return out;
}

string CudaLaneExpression(const CudaCoarsening& c, int real_extent, const string& lane) {
	static const char* const kAxis[] = {"x", "y", "z"};
	const string tid = string("(int)threadIdx.") + kAxis[c.dim];
	if (c.dim == 0) return tid + " + " + to_string(real_extent) + " * " + lane;  // a block apart: each load stays coalesced
	return tid + " * " + to_string(c.factor) + " + " + lane;                     // adjacent rows
}

string CoarsenBody(const string& body, const CudaCoarsening& c, int real_extent) {
	const vector<string> lines = SplitLines(body);
	const string tid_name = "block_thread_id" + to_string(c.dim);
	// locate the guard
	int guard = -1;
	for (int i = 0; i < (int)lines.size(); i++) {
		if (Trimmed(lines[i]).rfind("bool is_inside_dispatch = ", 0) == 0) {
			if (guard >= 0) return "";
			guard = i;
		}
	}
	if (guard < 0 || guard + 2 >= (int)lines.size()) return "";
	if (Trimmed(lines[guard + 1]) != "if (is_inside_dispatch)" || Trimmed(lines[guard + 2]) != "{") return "";
	int last = (int)lines.size() - 1;
	while (last >= 0 && Trimmed(lines[last]).empty()) last--;
	if (last <= guard + 2 || Trimmed(lines[last]) != "}") return "";
	vector<string> prologue, payload;
	for (int i = 0; i <= guard; i++) {
		const string t = Trimmed(lines[i]);
		if (t.empty()) continue;
		if (!IsPlainStatement(t)) return "";
		prologue.push_back(t);
	}
	// payload lines are statements (replicated per lane) or the structure of a loop every lane runs in step: `for (...)` with a header
	// that names nothing lane-dependent (the 9-tap loops of a filter), and its braces - emitted once around the lanes' statements
	vector<bool> shared;
	int depth = 0;
	for (int i = guard + 3; i < last; i++) {
		const string t = Trimmed(lines[i]);
		if (t.empty()) continue;
		const bool brace = t == "{" || t == "}";
		const bool loop_header = t.rfind("for (", 0) == 0 && t.back() == ')';
		if (!brace && !loop_header && !IsPlainStatement(t)) return "";
		if (t == "{") depth++;
		if (t == "}" && --depth < 0) return "";
		payload.push_back(t);
		shared.push_back(brace || loop_header);
	}
	if (depth != 0 || payload.empty() || payload.size() > 128) return "";
	// names that become per-lane: everything the body declares + the thread id the lanes run along
	unordered_set<string> names = {tid_name};
	unordered_set<string> all_tokens;
	for (const vector<string>* part : {&prologue, &payload}) {
		for (const string& t : *part) {
			const string name = DeclaredName(t);
			if (!name.empty()) names.insert(name);
			size_t i = 0;
			while (i < t.size()) {
				if (IsIdentStart(t[i])) {
					size_t j = i;
					while (j < t.size() && IsIdentChar(t[j])) j++;
					all_tokens.insert(t.substr(i, j - i));
					i = j;
				} else {
					i++;
				}
			}
		}
	}
	// a loop header must read the same for every lane
	for (size_t i = 0; i < payload.size(); i++) {
		if (!shared[i] || payload[i] == "{" || payload[i] == "}") continue;
		if (RenameIdentifiers(payload[i], names, "_") != payload[i]) return "";
	}
	auto suffix = [](int lane) { return "_L" + to_string(lane); };
	for (const string& name : names)
		for (int lane = 0; lane < c.factor; lane++)
			if (all_tokens.count(name + suffix(lane))) return "";  // a user name that already looks like a lane copy

	string out = "// " + to_string(c.factor) + " lanes per thread along block dimension " + to_string(c.dim) + " (thread coarsening)\n";
	for (int lane = 0; lane < c.factor; lane++)
		out += "int " + tid_name + suffix(lane) + " = " + CudaLaneExpression(c, real_extent, to_string(lane)) + ";\n";
	for (const string& t : prologue)
		for (int lane = 0; lane < c.factor; lane++) out += RenameIdentifiers(t, names, suffix(lane)) + "\n";
	out += "if (";
	for (int lane = 0; lane < c.factor; lane++) out += string(lane ? " && " : "") + "is_inside_dispatch" + suffix(lane);
	out += ")\n{\n";
	for (size_t i = 0; i < payload.size(); i++) {
		if (shared[i]) {
			out += "  " + payload[i] + "\n";
			continue;
		}
		for (int lane = 0; lane < c.factor; lane++) out += "  " + RenameIdentifiers(payload[i], names, suffix(lane)) + "\n";
	}
	out += "}\n";
	if (!c.exact) {
		// a block that crosses the edge of the dispatch: lane by lane through the untouched body
		out += "else\n{\n  #pragma unroll 1\n  for (int tf_lane = 0; tf_lane < " + to_string(c.factor) + "; tf_lane++)\n  {\n";
		out += "    int " + tid_name + " = " + CudaLaneExpression(c, real_extent, "tf_lane") + ";\n";
		for (const string& t : prologue) out += "    " + t + "\n";
		out += "    if (is_inside_dispatch)\n    {\n";
		for (const string& t : payload) out += "      " + t + "\n";
		out += "    }\n  }\n}\n";
	}
	return out;
}

// fallback when the body is not straight-line after all: the lanes one after the other around the untouched text.  Not usable when a
// lane can end the thread (`discard` is `return`) or meets a barrier.
string LaneLoopBody(const string& body, const CudaCoarsening& c, int real_extent) {
	if (body.find("discard") != string::npos || body.find("return") != string::npos || body.find("tf_group_barrier") != string::npos) return "";
	const string tid_name = "block_thread_id" + to_string(c.dim);
	string out = "#pragma unroll 1\nfor (int tf_lane = 0; tf_lane < " + to_string(c.factor) + "; tf_lane++)\n{\n";
	out += "  int " + tid_name + " = " + CudaLaneExpression(c, real_extent, "tf_lane") + ";\n";
	out += AddIndent(body, "  ");
	out += "}\n";
	return out;
}

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

	vector<int> group = kernel->root->group_size;  // the block the IR computed indices for ("virtual" when the kernel is coarsened)
	while (group.size() < 3) group.push_back(1);
	array<int, 3> launch_block = {group[0], group[1], group[2]};
	g_launch_block.erase(kernel->kernel_id_);

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
	string body = generator.AssembleString();

	// thread coarsening: the IR's block is `factor` times the launched one along c.dim; each thread carries `factor` lanes
	int coarsened_dim = -1;
	{
		auto it = g_coarsened.find(kernel->root);
		const bool group_memory = !GetGroupBufferDeclarations(kernel, CudaSharedDeclaration).empty();
		if (it != g_coarsened.end() && it->second.group == kernel->root->group_size && !group_memory) {
			const CudaCoarsening c = it->second;
			if (c.dim < 3 && c.factor > 1 && group[c.dim] % c.factor == 0) {
				const int real_extent = group[c.dim] / c.factor;
				// TFCUDA_COARSEN_FORCE_LOOP=1 (test aid): skip the replication so that the lane-loop fallback is exercised
				static const bool force_loop = getenv("TFCUDA_COARSEN_FORCE_LOOP") && atoi(getenv("TFCUDA_COARSEN_FORCE_LOOP")) != 0;
				string lanes = force_loop ? string() : CoarsenBody(body, c, real_extent);
				if (lanes.empty()) lanes = LaneLoopBody(body, c, real_extent);
				if (!lanes.empty()) {
					body = lanes;
					launch_block[c.dim] = real_extent;
					coarsened_dim = c.dim;
				}
				// otherwise the kernel is launched with the whole virtual block (at most 1024 threads by construction)
			}
		}
	}
	g_launch_block[kernel->kernel_id_] = launch_block;
	const int threads = launch_block[0] * launch_block[1] * launch_block[2];

	string bindings = "struct " + args_t + " {\n";
	if (n_mem > 0) bindings += "  uint* mem[" + to_string(n_mem) + "];\n";
	bindings += "  uint var[" + to_string(n_var) + "];\n};\n";

	// TFCUDA_MIN_BLOCKS=n (tuning aid): ask for n resident blocks per SM, i.e. cap the registers per thread at 65536 / (n * threads)
	string bounds = to_string(threads);
	if (const char* mb = getenv("TFCUDA_MIN_BLOCKS")) {
		if (atoi(mb) > 0) bounds += ", " + to_string(atoi(mb));
	}
	// the block the kernel must be launched with, for everything that launches emitted text without this backend (tests/cpu_sim,
	// tests/standalone): the host program's tf.dispatch line carries the IR's block, which differs for coarsened kernels
	string main_code;
	main_code = "// tfcuda_block: " + to_string(launch_block[0]) + " " + to_string(launch_block[1]) + " " + to_string(launch_block[2]) + "\n";
	main_code += "extern \"C\" __global__ void __launch_bounds__(" + bounds + ") " + kname +
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
	static const char* const kAxis[] = {"x", "y", "z"};
	main_code += "  (void)block_id;\n";
	for (int d = 0; d < 3; d++) {
		if (d == coarsened_dim) continue;  // declared per lane by the body
		main_code += "  int block_thread_id" + to_string(d) + " = (int)threadIdx." + kAxis[d] + "; (void)block_thread_id" + to_string(d) + ";\n";
	}
	main_code += "\n";
	main_code += AddIndent(body, "  ");
	main_code += "}\n";

	kernel->generated_header_ = "";  // the prelude is shared by all kernels: libtfcuda prepends it per NVRTC unit
	kernel->generated_bindings_ = bindings;
	kernel->generated_main_ = main_code;
	kernel->full_generated_code_ = bindings + main_code;
}

}  // namespace TensorFrost
