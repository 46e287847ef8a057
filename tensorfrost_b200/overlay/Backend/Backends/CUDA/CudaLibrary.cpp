// Attaches the hand-written sm_100a library kernels of libtfcuda.so to COMPILED TensorFrost programs.
//
// The reference lowers its algorithmic ops (`matmul`, `dim_sum/max/min/mean/norm`, `dim_prefix_sum`) to per-output-element
// serial loops through the `implementation_functions` table (Compiler/Implementations.cpp:648-728), consulted by
// IR::InsertAlgorithmicPrimitives (Compiler/Steps/Algorithms.cpp:4-62).  When the CUDA backend is initialised this file
// replaces those table entries.  A replacement does not emit a loop; it emits a *library call*:
//
//     out = memory(shape of the op's result)
//     kernel([1]) "tfcuda_lib:<op>:<params>" { store(out, f(load(in0), load(in1), ...)) }     <- never executed as such
//     result = load(out, element indices)                                                    <- fuses into the consumers
//
// The one-thread kernel exists so that the UNCHANGED compiler does all the bookkeeping: inputs are materialised (StopFusion),
// GenerateProgram collects the memory bindings (Compiler/KernelGen.cpp:13-68), the host program gets an ordinary
// `tf.dispatch(id, {out}, {in...}, ...)` (CodeGen/Langs/CPP.cpp:586-637) and allocation / deallocation of `out` is planned as
// for any buffer.  The CUDA emitter (CodeGen/Langs/CUDA.cpp) recognises the marker, records which binding plays which role and
// emits no source; CudaKernelManager::DispatchKernel then calls tfcuda_matmul / tfcuda_reduce / tfcuda_prefix_sum /
// tfcuda_radix_sort with the device pointers and the extents taken from the TFTensor shapes of the dispatch.
// Autodiff is unaffected: VJPs are defined on the un-lowered ops (Implementations.cpp:133-172) and run before this lowering.
//
// Policy (when the generic fused loop is kept instead): see WantsLibraryReduction / the matmul rules below.  TFCUDA_LIBRARY=0
// in the environment keeps the reference's generic lowering everywhere (used by the parity tests to cover both paths).
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "CUDA.h"
#include "Compiler/Implementations.h"
#include "Backend/Backend.h"

#define TFCUDA_NO_ABI_STRUCTS
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

extern map<string, ImplementationFunction> implementation_functions;  // Compiler/Implementations.cpp:648
extern map<string, VJPGradientFunction> gradient_functions;           // Compiler/Implementations.cpp:86

namespace {

const char* const kMarker = "tfcuda_lib:";
map<string, ImplementationFunction> g_generic;        // the reference's own lowerings, kept for the cases the library declines
VJPGradientFunction g_generic_matmul_vjp;             // the reference's matmul VJP (Implementations.cpp:133-135)
unordered_map<size_t, CudaLibraryCall> g_calls;       // kernel id -> library call (filled by the emitter)
bool g_installed = false;

int EnvInt(const char* name, int fallback) {
	const char* v = getenv(name);
	return v ? atoi(v) : fallback;
}

bool LibraryEnabled() { return current_kernel_lang == CodeGenLang::CUDA && EnvInt("TFCUDA_LIBRARY", 1) != 0; }

bool IsStoredTensor(const Tensor* t) { return t->node_->name == "memory"; }  // program inputs and tf.buffer tensors

const map<string, int> kReduceOps = {{"dim_sum", TFCUDA_RED_SUM}, {"dim_max", TFCUDA_RED_MAX}, {"dim_min", TFCUDA_RED_MIN},
                                     {"dim_mean", TFCUDA_RED_MEAN}, {"dim_norm", TFCUDA_RED_NORM}};

// A reduction / scan goes to the library when the reduced axis is long enough for a cooperative kernel to win over the fused
// serial loop: a compile-time extent >= 1024 (the size at which the reference itself gives up on one loop and stages the
// reduction, Steps/Optimization.cpp:469-470), or a run-time extent over a tensor that is already stored (nothing to re-materialise).
bool WantsLibraryAxis(const Tensor* input, int axis) {
	Tensors shape = input->GetShape();
	int dims = (int)shape.size();
	if (axis < 0) axis += dims;
	if (axis < 0 || axis >= dims) return false;
	TFType type = input->node_->format.type;
	if (type != TFType::Float && type != TFType::Int && type != TFType::Uint) return false;
	int extent = shape[axis]->TryGetConstant();
	if (extent >= EnvInt("TFCUDA_LIBRARY_MIN_AXIS", 1024)) return true;
	return extent < 0 && IsStoredTensor(input);
}

Tensors ZeroIndices(int dims) {
	Tensors idx;
	for (int d = 0; d < dims; d++) idx.push_back(&Tensor::Constant(0));
	return idx;
}

struct OutputSpec {
	Tensors shape;
	TFDataFormat format;
};

// Emits the marker kernel; returns the output buffers (memory tensors).
vector<Tensor*> EmitLibraryCall(const string& marker, const vector<const Tensor*>& inputs, const vector<OutputSpec>& outputs) {
	for (const Tensor* in : inputs) in->StopFusion();
	vector<Tensor*> buffers;
	for (const OutputSpec& o : outputs) {
		Tensor& buf = Tensor::Memory(o.shape, o.format);
		buf.SetDebugName("lib_out");
		buffers.push_back(&buf);
	}
	Tensor& kernel = Tensor::Kernel({&Tensor::Constant(1)}, [&](Tensors) {
		const Tensor* mix = nullptr;
		for (const Tensor* in : inputs) {
			Tensor& value = Tensor::Load(*in, ZeroIndices(in->GetDimension()), IndexingMode::Unsafe);
			value.node_->flags.set(NodeProp::NoLoadFusion);
			const Tensor* bits = &Tensor::asuint(value);
			mix = mix ? &(*mix ^ *bits) : bits;
		}
		if (mix == nullptr) mix = &Tensor::Constant(0u);
		for (size_t i = 0; i < buffers.size(); i++) {
			const Tensor* value = mix;
			if (outputs[i].format.type == TFType::Float) value = &Tensor::asfloat(*mix);
			else if (outputs[i].format.type == TFType::Int) value = &Tensor::asint(*mix);
			Tensor::Store(*buffers[i], *value, ZeroIndices((int)outputs[i].shape.size()), IndexingMode::Unsafe);
		}
	}, {1});
	kernel.node_->debug_name = marker;  // set directly: survives tf.strip_debug_info()
	return buffers;
}

Tensor* ElementView(Tensor* buffer, const Tensors& shape) {
	Tensors idx;
	for (int i = 0; i < (int)shape.size(); i++) idx.push_back(&Tensor::Index(shape, i));
	return &Tensor::Load(*buffer, idx, IndexingMode::Unsafe);
}

// ---- lowerings ---------------------------------------------------------------------------------------------------
void LowerReduction(const string& name, Tensors& outputs, map<int, const Tensor*> inputs, const Tensor* tensor, vector<int> axes) {
	const Tensor* in = inputs[0];
	int dims = in->GetDimension();
	int axis = axes[0] < 0 ? axes[0] + dims : axes[0];
	if (!LibraryEnabled() || !WantsLibraryAxis(in, axis)) {
		g_generic[name](outputs, inputs, tensor, axes);
		return;
	}
	Tensors out_shape = tensor->GetShape();
	string marker = string(kMarker) + "reduce:" + to_string(kReduceOps.at(name)) + ":" + to_string(axis);
	vector<Tensor*> bufs = EmitLibraryCall(marker, {in}, {{out_shape, tensor->node_->format}});
	Tensor* result = ElementView(bufs[0], out_shape);
	result->SetDebugName(name.substr(4));
	outputs.push_back(result);
}

void LowerPrefixSum(Tensors& outputs, map<int, const Tensor*> inputs, const Tensor* tensor, vector<int> axes) {
	const Tensor* in = inputs[0];
	int dims = in->GetDimension();
	int axis = axes[0] < 0 ? axes[0] + dims : axes[0];
	if (!LibraryEnabled() || !WantsLibraryAxis(in, axis)) {
		g_generic["dim_prefix_sum"](outputs, inputs, tensor, axes);
		return;
	}
	Tensors out_shape = tensor->GetShape();
	vector<Tensor*> bufs = EmitLibraryCall(string(kMarker) + "scan:" + to_string(axis), {in}, {{out_shape, tensor->node_->format}});
	Tensor* result = ElementView(bufs[0], out_shape);
	result->SetDebugName("prefix_sum");
	outputs.push_back(result);
}

// Is `t` the lowered form of a plain 2-D transpose (ComputeTranspose, Implementations.cpp:455-489: an Unsafe load of the source
// with the two dim_id indices swapped)?  Returns the source tensor S (t = S^T) or nullptr.
const Tensor* TransposedSource(const Tensor* t) {
	Node* node = t->node_;
	if (node->name != "load" || t->GetDimension() != 2) return nullptr;
	if (!node->args.Has(ArgType::Memory) || !node->args.Has(ArgType::Index, 0) || !node->args.Has(ArgType::Index, 1)) return nullptr;
	if (node->args.Has(ArgType::Index, 2)) return nullptr;
	const Tensor* source = node->args.Get(ArgType::Memory)->GetTensor();
	if (source->GetDimension() != 2) return nullptr;
	Node* i0 = node->args.Get(ArgType::Index, 0);
	Node* i1 = node->args.Get(ArgType::Index, 1);
	if (i0->name != "dim_id" || i1->name != "dim_id") return nullptr;
	if ((int)i0->data[0] != 1 || (int)i1->data[0] != 0) return nullptr;
	// the load must cover the WHOLE source (a cropped or broadcast transpose is not S^T): extents equal as constants or as the same node
	Tensors ts = t->GetShape(), ss = source->GetShape();
	for (int d = 0; d < 2; d++) {
		const Tensor* x = ts[d];
		const Tensor* y = ss[1 - d];
		int cx = x->TryGetConstant(), cy = y->TryGetConstant();
		bool same = (x == y) || (x->node_ == y->node_) || (cx >= 0 && cx == cy);
		if (!same) return nullptr;
	}
	return source;
}

// matmul: A [.., M, K] @ B [K, N] (B 2-D: the leading dims of A fold into M) or equal-rank batched operands; fp32 only.
void LowerMatmul(Tensors& outputs, map<int, const Tensor*> inputs, const Tensor* tensor, vector<int> axes) {
	const Tensor* a = inputs[0];
	const Tensor* b = inputs[1];
	int da = a->GetDimension(), db = b->GetDimension();
	// batched operands (B with batch dims, e.g. the per-sample products of matmul's own VJP) stay on the generic lowering: one library
	// launch per batch slice would be launch-bound for the many-small-slices shapes autodiff produces
	bool shapes_ok = da >= 2 && db == 2;
	bool types_ok = a->node_->format.type == TFType::Float && b->node_->format.type == TFType::Float;
	if (!LibraryEnabled() || !shapes_ok || !types_ok || EnvInt("TFCUDA_LIBRARY_MATMUL", 1) == 0) {
		g_generic["matmul"](outputs, inputs, tensor, axes);
		return;
	}
	Tensors out_shape = tensor->GetShape();
	// S^T @ B with 2-D S [R, M] and B [R, N] (what CudaMatmulVJP emits for weight gradients, or a user's `x.T @ y`): contracted over
	// the leading extent of both operands as they lie in memory, no transposed copy (tfcuda_matmul_tn)
	if (da == 2 && EnvInt("TFCUDA_LIBRARY_MATMUL_TN", 1) != 0) {
		if (const Tensor* source = TransposedSource(a)) {
			vector<Tensor*> bufs = EmitLibraryCall(string(kMarker) + "matmul_tn", {source, b}, {{out_shape, tensor->node_->format}});
			Tensor* result = ElementView(bufs[0], out_shape);
			result->SetDebugName("matmul_tn");
			outputs.push_back(result);
			return;
		}
	}
	// precision of the product: tf.initialize(tf.cuda, "--tf-matmul=tf32|3xtf32|fp32"), or TFCUDA_MATMUL_MODE=0|1|2 in the environment;
	// default 3xTF32 (fp32-accurate: north_star's 1e-5 class); tf32 is inside north_star's 1e-3 matmul bar and ~3x faster
	int mode = EnvInt("TFCUDA_MATMUL_MODE", 1);
	const std::string opt = CudaBackendOption("matmul", "");
	if (opt == "tf32") mode = 0;
	else if (opt == "3xtf32") mode = 1;
	else if (opt == "fp32") mode = 2;
	else if (!opt.empty()) throw std::runtime_error("CUDA backend: --tf-matmul must be tf32, 3xtf32 or fp32");
	vector<Tensor*> bufs = EmitLibraryCall(string(kMarker) + "matmul:" + to_string(mode), {a, b}, {{out_shape, tensor->node_->format}});
	Tensor* result = ElementView(bufs[0], out_shape);
	result->SetDebugName("matmul");
	outputs.push_back(result);
}

// VJP of C = A @ B on this backend.  dA = dC @ B^T as in the reference (Implementations.cpp:133-135).  For a 2-D B (a weight matrix
// applied to every row of an N-D A) the reference forms dB = Transpose(A)[batch] @ dC[batch] per leading slice and lets
// ReduceGradientToShape sum the [batch.., K, N] products over the batch axes (:5-54).  The same value is ONE contraction over all rows:
// dB = A2^T @ dC2 with A2 = A viewed as [rows, K] and dC2 = dC viewed as [rows, N]; LowerMatmul sends it to tfcuda_matmul_tn.
void CudaMatmulVJP(ArgumentManager& in, const Tensor& out, const Tensor& grad, NodeGrads& grads) {
	const Tensor& a = in[0];
	const Tensor& b = in[1];
	bool fold = LibraryEnabled() && EnvInt("TFCUDA_LIBRARY_MATMUL", 1) != 0 && EnvInt("TFCUDA_LIBRARY_MATMUL_TN", 1) != 0 &&
	            b.GetDimension() == 2 && a.GetDimension() > 2 && a.node_->format.type == TFType::Float && b.node_->format.type == TFType::Float;
	if (!fold) {
		g_generic_matmul_vjp(in, out, grad, grads);
		return;
	}
	// internal shapes are innermost-first: A = [K, M, batch...], dC = [N, M, batch...]
	Tensors sa = a.GetShape();
	Tensors sg = grad.GetShape();
	const Tensor* rows = sa[1];
	long long const_rows = sa[1]->TryGetConstant();
	for (size_t i = 2; i < sa.size(); i++) {
		int c = sa[i]->TryGetConstant();
		const_rows = (const_rows >= 0 && c >= 0) ? const_rows * c : -1;
		rows = &(*rows * *sa[i]);
	}
	if (const_rows >= 0 && const_rows < 0x7fffffffLL) rows = &Tensor::Constant((int)const_rows);
	Tensor& a2 = Tensor::Reshape(a, {sa[0], rows});
	Tensor& g2 = Tensor::Reshape(grad, {sg[0], rows});
	grads.Add(Tensor::Matmul(grad, Tensor::Transpose(b)), Tensor::Matmul(Tensor::Transpose(a2), g2));
}

size_t Extent(const TFTensor& t, size_t from, size_t to) {
	size_t n = 1;
	for (size_t i = from; i < to; i++) n *= t.shape[i];
	return n;
}

uint64_t Ptr(const TFTensor& t) { return ((TFCudaBuffer*)t.buffer)->GetNative(); }

void Check(int rc, const string& what) {
	if (rc != 0) throw std::runtime_error("CUDA backend: library call " + what + " failed: " + tfcuda_last_error());
}

}  // namespace

bool CudaLibraryWantsReduction(Node* node) {
	if (!LibraryEnabled() || !g_installed || !kReduceOps.contains(node->name)) return false;
	const Tensor* input = node->args.Get(ArgType::Input, 0)->GetTensor();
	return WantsLibraryAxis(input, (int)node->data[0]);
}

void InstallCudaLibraryLowerings() {
	if (g_installed) return;
	for (auto& [name, op] : kReduceOps) {
		g_generic[name] = implementation_functions.at(name);
		string n = name;
		implementation_functions[name] = [n](Tensors& out, map<int, const Tensor*> in, const Tensor* t, vector<int> axes) { LowerReduction(n, out, in, t, axes); };
	}
	g_generic["dim_prefix_sum"] = implementation_functions.at("dim_prefix_sum");
	implementation_functions["dim_prefix_sum"] = LowerPrefixSum;
	g_generic["matmul"] = implementation_functions.at("matmul");
	implementation_functions["matmul"] = LowerMatmul;
	g_generic_matmul_vjp = gradient_functions.at("matmul");
	gradient_functions["matmul"] = CudaMatmulVJP;
	g_installed = true;
}

// The traced form of tf.sort.radix on this backend (called from the python binding): one library call, stable LSD radix.
vector<Tensor*> CudaLibrarySort(const Tensor* keys, const Tensor* values, int max_bits) {
	if (keys->GetDimension() != 1) throw std::runtime_error("cuda radix sort: keys must be one-dimensional");
	Tensors shape = keys->GetShape();
	// the sort's scratch is NOT a program buffer: the compiler prunes memory that is never loaded; DispatchCudaLibraryCall takes it
	// from the runtime's stream-ordered pool for the duration of the call instead
	vector<const Tensor*> inputs = {keys};
	vector<OutputSpec> outputs = {{shape, keys->node_->format}};
	if (values != nullptr) {
		inputs.push_back(values);
		outputs.push_back({shape, values->node_->format});
	}
	string marker = string(kMarker) + "sort:" + to_string(values != nullptr ? 1 : 0) + ":" + to_string(max_bits);
	vector<Tensor*> bufs = EmitLibraryCall(marker, inputs, outputs);
	vector<Tensor*> result;
	result.push_back(ElementView(bufs[0], shape));
	result[0]->SetDebugName("sorted_keys");
	if (values != nullptr) {
		result.push_back(ElementView(bufs[1], shape));
		result[1]->SetDebugName("sorted_values");
	}
	return result;
}

// ---- emitter side ------------------------------------------------------------------------------------------------
bool IsCudaLibraryKernel(Kernel* kernel) { return kernel->root->debug_name.rfind(kMarker, 0) == 0; }

void RegisterCudaLibraryKernel(Kernel* kernel) {
	CudaLibraryCall call;
	std::stringstream ss(kernel->root->debug_name.substr(strlen(kMarker)));
	string field;
	std::getline(ss, call.op, ':');
	while (std::getline(ss, field, ':')) call.params.push_back(atoi(field.c_str()));
	map<Node*, size_t> bindings = kernel->GetMemoryBindings();
	for (auto node = NodeIterator(kernel->root); !node.end(); node.next()) {
		if (node->name != "load" && node->name != "store") continue;
		if (node->flags.has(NodeProp::LocalMemoryOp)) continue;
		Node* memory = node->args.GetTensor(ArgType::Memory)->node_;
		if (!bindings.contains(memory)) throw std::runtime_error("CUDA emitter: library call " + call.op + " refers to an unbound tensor");
		(node->name == "load" ? call.inputs : call.outputs).push_back((int)bindings[memory]);
	}
	g_calls[kernel->kernel_id_] = call;
}

const CudaLibraryCall* FindCudaLibraryCall(size_t kernel_id) {
	auto it = g_calls.find(kernel_id);
	return it == g_calls.end() ? nullptr : &it->second;
}

// ---- dispatch side -----------------------------------------------------------------------------------------------
void DispatchCudaLibraryCall(const CudaLibraryCall& call, const TFDispatchInfo& info) {
	auto tensor = [&](int binding) -> const TFTensor& {
		if (binding < 0 || (size_t)binding >= info.read_write_count) throw std::runtime_error("CUDA backend: library call binding out of range");
		return info.read_write_tensors[binding];
	};
	auto need = [&](size_t ins, size_t outs, size_t params) {
		if (call.inputs.size() != ins || call.outputs.size() != outs || call.params.size() < params)
			throw std::runtime_error("CUDA backend: malformed library call " + call.op + " (" + to_string(call.inputs.size()) + " inputs, " +
			                         to_string(call.outputs.size()) + " outputs)");
	};
	if (call.op == "reduce" || call.op == "scan") {
		bool reduce = call.op == "reduce";
		need(1, 1, reduce ? 2 : 1);
		const TFTensor& in = tensor(call.inputs[0]);
		const TFTensor& out = tensor(call.outputs[0]);
		int internal_axis = reduce ? call.params[1] : call.params[0];
		size_t axis = in.dim - 1 - (size_t)internal_axis;  // IR dims are innermost-first, TFTensor shapes outermost-first
		size_t outer = Extent(in, 0, axis), n = in.shape[axis], inner = Extent(in, axis + 1, in.dim);
		if (outer * n * inner == 0) return;
		if (reduce) Check(tfcuda_reduce(Ptr(in), Ptr(out), outer, n, inner, call.params[0], (int)in.format.type), "reduce");
		else Check(tfcuda_prefix_sum(Ptr(in), Ptr(out), outer, n, inner, (int)in.format.type), "prefix_sum");
	} else if (call.op == "matmul") {
		need(2, 1, 1);
		const TFTensor& a = tensor(call.inputs[0]);
		const TFTensor& b = tensor(call.inputs[1]);
		const TFTensor& c = tensor(call.outputs[0]);
		size_t k = a.shape[a.dim - 1], n = b.shape[b.dim - 1];
		if (b.shape[b.dim - 2] != k) throw std::runtime_error("CUDA backend: matmul inner dimensions differ at run time");
		size_t batch = 1, m;
		if (b.dim == 2) {
			m = Extent(a, 0, a.dim - 1);
		} else {
			batch = Extent(a, 0, a.dim - 2);
			if (batch != Extent(b, 0, b.dim - 2)) throw std::runtime_error("CUDA backend: batched matmul needs equal batch extents");
			m = a.shape[a.dim - 2];
		}
		if (batch * m * n == 0) return;
		// experimental (round 1: compiled, not yet validated on hardware): small weight matrix applied to very many rows
		static const bool rows_kernel = EnvInt("TFCUDA_MATMUL_ROWS", 0) != 0;
		if (rows_kernel && batch == 1 && m >= 4096 && tfcuda_matmul_rows_supported(m, k, n)) {
			Check(tfcuda_matmul_rows(Ptr(a), Ptr(b), Ptr(c), m, k, n), "matmul_rows");
			return;
		}
		Check(tfcuda_matmul(Ptr(a), Ptr(b), Ptr(c), batch, m, n, k, call.params[0]), "matmul");
	} else if (call.op == "matmul_tn") {
		need(2, 1, 0);
		const TFTensor& a = tensor(call.inputs[0]);  // [R, M] (possibly a reshaped view of an N-D tensor)
		const TFTensor& b = tensor(call.inputs[1]);  // [R, N]
		const TFTensor& c = tensor(call.outputs[0]);
		size_t m = a.shape[a.dim - 1], n = b.shape[b.dim - 1];
		size_t r = Extent(a, 0, a.dim - 1);
		if (Extent(b, 0, b.dim - 1) != r) throw std::runtime_error("CUDA backend: matmul_tn operands have different row counts at run time");
		if (Extent(c, 0, c.dim) != m * n) throw std::runtime_error("CUDA backend: matmul_tn output extent mismatch");
		if (m * n == 0) return;
		Check(tfcuda_matmul_tn(Ptr(a), Ptr(b), Ptr(c), r, m, n), "matmul_tn");
	} else if (call.op == "sort") {
		bool has_values = call.params.size() > 0 && call.params[0] != 0;
		need(has_values ? 2 : 1, has_values ? 2 : 1, 2);
		const TFTensor& keys = tensor(call.inputs[0]);
		size_t n = Extent(keys, 0, keys.dim);
		if (n == 0) return;
		uint64_t temp = tfcuda_malloc(tfcuda_radix_sort_temp_words(n) * 4);  // cudaMallocAsync on the runtime stream: pool hit after warm-up
		if (!temp) throw std::runtime_error(string("CUDA backend: radix sort scratch: ") + tfcuda_last_error());
		int rc = tfcuda_radix_sort(Ptr(keys), Ptr(tensor(call.outputs[0])), has_values ? Ptr(tensor(call.inputs[1])) : 0,
		                           has_values ? Ptr(tensor(call.outputs[1])) : 0, n, (int)keys.format.type, call.params[1], temp);
		tfcuda_free(temp);  // stream-ordered: released after the sort's kernels
		Check(rc, "radix_sort");
	} else {
		throw std::runtime_error("CUDA backend: unknown library call " + call.op);
	}
}

}  // namespace TensorFrost
