// Python surface of the CUDA backend: `CudaDefinitions(m)` is called from PYBIND11_MODULE
// (Frontend/Python/PybindModule.cpp:96-101 calls the other *Definitions the same way).
//
// Nothing here changes an existing tf.* function.  It adds what a device backend needs next to them:
//   tf.cuda_synchronize / cuda_timer_* / cuda_launch_count   — device-side timing for benchmarks
//   tf.cuda_tensor(np) / tf.cuda_numpy(t) / tf.cuda_upload / tf.cuda_download — bulk host<->device copies that skip the
//        per-element std::function conversion of PyTensorMemory (Frontend/Python/PyTensorMemory.cpp:19-84,
//        PyTensorMemory.h:30-45); same dtype rules for the 4-byte types, same result objects
//   tf.cuda_device_ptr(t)                                     — interop (e.g. torch via __cuda_array_interface__)
//   tf.cuda_comm_* / tf.cuda_allreduce                        — the data-parallel gradient exchange
//   tf.cuda_radix_sort / cuda_reduce / cuda_prefix_sum / cuda_matmul / cuda_scatter_add / cuda_nbody_step
//        — the hand-written library kernels on TensorMemory objects
#include <Frontend/Python/PyTensor.h>
#include <Frontend/Python/PyTensorMemory.h>

#include <deque>
#include <tuple>

#include "CUDA.h"

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


// ---- bulk paths behind the reference's own tf.tensor(np) / .numpy (overlay insertions in Frontend/Python/PyTensorMemory.{h,cpp}) ----
// tf.tensor(array): a C-contiguous 4-byte array goes to the device with ONE copy instead of the per-element std::function walk of
// PyTensorMemory.cpp:15-84 (seconds at 2^28 elements).  Anything else (strided views, 8-byte dtypes, bool) returns false and takes
// the reference path unchanged.
bool CudaBulkTensor(const py::buffer_info& info, TFTensor** out) {
	if (current_backend != BackendType::CUDA || info.itemsize != 4 || info.size <= 0 || info.ndim < 1) return false;
	TFDataFormat fmt;
	switch (info.format[0]) {
		case 'f': fmt = TFTypeFloat32; break;
		case 'i': case 'l': fmt = TFTypeInt32; break;
		case 'I': case 'L': fmt = TFTypeUint32; break;
		default: return false;
	}
	py::ssize_t expect = 4;
	for (py::ssize_t d = info.ndim - 1; d >= 0; d--) {
		if (info.shape[d] != 1 && info.strides[d] != expect) return false;
		expect *= info.shape[d];
	}
	std::vector<size_t> shape(info.shape.begin(), info.shape.end());
	TFTensor* t = global_memory_manager->AllocateTensor(shape, fmt);
	if (tfcuda_memcpy_h2d(((TFCudaBuffer*)t->buffer)->GetNative(), info.ptr, (size_t)info.size * 4) != 0)
		throw std::runtime_error(std::string("tf.tensor: upload failed: ") + tfcuda_last_error());
	*out = t;
	return true;
}

// .numpy: device -> the numpy array's own storage, no intermediate std::vector (PyTensorMemory.h:30-45)
bool CudaBulkReadback(const TFTensor* t, void* dst, size_t item_size) {
	if (current_backend != BackendType::CUDA || item_size != 4) return false;
	size_t n = GetSize(t);
	if (n == 0) return true;
	if (tfcuda_memcpy_d2h(dst, ((TFCudaBuffer*)t->buffer)->GetNative(), n * 4) != 0)
		throw std::runtime_error(std::string(".numpy: download failed: ") + tfcuda_last_error());
	return true;
}

namespace {

// ---- DLPack (v0.8 ABI, the subset needed for export) -----------------------------------------------------------------
struct DLDevice { int32_t device_type; int32_t device_id; };
struct DLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor { void* data; DLDevice device; int32_t ndim; DLDataType dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset; };
struct DLManagedTensor { DLTensor dl_tensor; void* manager_ctx; void (*deleter)(DLManagedTensor*); };
constexpr int32_t kDLCUDA = 2;
struct DLOwner { PyObject* tensor; std::vector<int64_t> shape; DLManagedTensor managed; };

const char* TypeStr(TFDataFormat f) {
	if (f == TFTypeFloat32) return "<f4";
	if (f == TFTypeInt32) return "<i4";
	if (f == TFTypeUint32) return "<u4";
	throw std::runtime_error("device interop: only float32 / int32 / uint32 tensors are exported");
}

int CurrentDevice() { return tfcuda_device_index(); }

}  // namespace

namespace {

void RequireCuda(const char* what) {
	if (current_backend != BackendType::CUDA) {
		throw std::runtime_error(std::string(what) + " requires tf.initialize(tf.cuda)");
	}
}

void Check(int rc, const char* what) {
	if (rc != 0) throw std::runtime_error(std::string(what) + ": " + tfcuda_last_error());
}

uint64_t DevPtr(const PyTensorMemory& t) { return ((TFCudaBuffer*)t.tensor_->buffer)->GetNative(); }

TFDataFormat FormatOf(const py::buffer_info& info) {
	if (info.itemsize != 4) throw std::runtime_error("cuda_tensor: only 4-byte dtypes (float32/int32/uint32) take the bulk path");
	switch (info.format[0]) {
		case 'f': return TFTypeFloat32;
		case 'i': case 'l': return TFTypeInt32;
		case 'I': case 'L': return TFTypeUint32;
		default: throw std::runtime_error("cuda_tensor: unsupported dtype format " + info.format);
	}
}

PyTensorMemory* NewLike(const PyTensorMemory& t, TFDataFormat fmt) {
	return new PyTensorMemory(global_memory_manager->AllocateTensor(t.Shape(), fmt));
}

}  // namespace

void CudaDefinitions(py::module& m) {
	m.def("cuda_synchronize", []() { RequireCuda("cuda_synchronize"); CudaFinish(); }, "Wait for all queued device work");
	m.def("cuda_launch_count", []() { return tfcuda_launch_count(); }, "Kernels launched since initialisation");
	m.def("cuda_pool_driver_calls", []() { return tfcuda_pool_driver_calls(); },
	      "cudaMallocAsync + cudaFreeAsync calls made for tensor buffers since initialisation (parked blocks are reused without one)");
	m.def("cuda_timer_begin", []() { Check(tfcuda_timer_begin(), "cuda_timer_begin"); });
	m.def("cuda_timer_end", []() { float ms = 0; Check(tfcuda_timer_end(&ms), "cuda_timer_end"); return ms; },
	      "Milliseconds of device time since cuda_timer_begin (CUDA events on the backend stream)");
	m.def("cuda_device_name", []() { return std::string(tfcuda_device_name()); });
	m.def("cuda_sm_count", []() { return tfcuda_device_sm_count(); });
	m.def("cuda_device_ptr", [](const PyTensorMemory& t) { RequireCuda("cuda_device_ptr"); return DevPtr(t); });

	m.def("cuda_profile_enable", [](bool on) { Check(tfcuda_profile_enable(on ? 1 : 0), "cuda_profile_enable"); }, py::arg("on") = true,
	      "Bracket every kernel launch with CUDA events and accumulate per-kernel time / algorithmic bytes");
	m.def("cuda_profile_reset", []() { tfcuda_profile_reset(); });
	m.def("cuda_profile_records", []() {
		size_t n = tfcuda_profile_records(nullptr, 0);
		std::vector<TFCudaProfileRecord> recs(n);
		n = tfcuda_profile_records(recs.data(), recs.size());
		py::list out;
		for (size_t i = 0; i < n && i < recs.size(); i++) {
			py::dict d;
			d["name"] = std::string(recs[i].name);
			d["launches"] = recs[i].launches;
			d["total_ms"] = recs[i].total_ms;
			d["bytes"] = recs[i].bytes;
			out.append(d);
		}
		return out;
	}, "Per-kernel profile records: name, launches, total_ms (CUDA events), bytes (tensors bound, each once)");

	m.def("cuda_pinned_array", [](std::vector<size_t> shape, const std::string& dtype) -> py::array {
		size_t count = 1;
		for (size_t d : shape) count *= d;
		void* p = tfcuda_host_alloc(count * 4);
		if (!p) throw std::runtime_error(std::string("cuda_pinned_array: ") + tfcuda_last_error());
		py::capsule owner(p, [](void* q) { tfcuda_host_free(q); });
		if (dtype == "float32") return py::array_t<float>(shape, static_cast<float*>(p), owner);
		if (dtype == "int32") return py::array_t<int>(shape, static_cast<int*>(p), owner);
		if (dtype == "uint32") return py::array_t<uint>(shape, static_cast<uint*>(p), owner);
		throw std::runtime_error("cuda_pinned_array: dtype must be float32, int32 or uint32");
	}, py::arg("shape"), py::arg("dtype") = "float32", "numpy array backed by page-locked host memory (full-rate host<->device copies)");

	m.def("cuda_tensor", [](py::array arr) {
		RequireCuda("cuda_tensor");
		py::array c = py::array::ensure(arr, py::array::c_style);
		py::buffer_info info = c.request();
		TFDataFormat fmt = FormatOf(info);
		std::vector<size_t> shape(info.shape.begin(), info.shape.end());
		TFTensor* t = global_memory_manager->AllocateTensor(shape, fmt);
		Check(tfcuda_memcpy_h2d(((TFCudaBuffer*)t->buffer)->GetNative(), info.ptr, (size_t)info.size * 4), "cuda_tensor upload");
		return new PyTensorMemory(t);
	}, "TensorMemory from a contiguous 4-byte numpy array with one bulk copy", py::return_value_policy::take_ownership);

	m.def("cuda_upload", [](PyTensorMemory& t, py::array arr) {
		RequireCuda("cuda_upload");
		py::array c = py::array::ensure(arr, py::array::c_style);
		py::buffer_info info = c.request();
		(void)FormatOf(info);
		if ((size_t)info.size != GetSize(t.tensor_)) throw std::runtime_error("cuda_upload: element count mismatch");
		Check(tfcuda_memcpy_h2d(DevPtr(t), info.ptr, (size_t)info.size * 4), "cuda_upload");
	}, "Overwrite an existing TensorMemory from a numpy array of the same size");

	m.def("cuda_download", [](const PyTensorMemory& t, py::array arr) {
		RequireCuda("cuda_download");
		py::buffer_info info = arr.request(true);  // writable
		(void)FormatOf(info);
		if (!(arr.flags() & py::array::c_style)) throw std::runtime_error("cuda_download: the destination array must be C-contiguous");
		if ((size_t)info.size != GetSize(t.tensor_)) throw std::runtime_error("cuda_download: element count mismatch");
		Check(tfcuda_memcpy_d2h(info.ptr, DevPtr(t), (size_t)info.size * 4), "cuda_download");
	}, "Copy a TensorMemory into an existing numpy array of the same size (a tf.cuda_pinned_array destination runs at full PCIe rate)");

	// copy engines: uploads / downloads on their own streams (see include/tfcuda.h); arrays should come from tf.cuda_pinned_array
	m.def("cuda_upload_async", [](PyTensorMemory& t, py::array arr) {
		RequireCuda("cuda_upload_async");
		py::buffer_info info = arr.request();
		(void)FormatOf(info);
		if (!(arr.flags() & py::array::c_style)) throw std::runtime_error("cuda_upload_async: the source array must be C-contiguous");
		if ((size_t)info.size != GetSize(t.tensor_)) throw std::runtime_error("cuda_upload_async: element count mismatch");
		Check(tfcuda_memcpy_h2d_async(DevPtr(t), info.ptr, (size_t)info.size * 4), "cuda_upload_async");
	}, "Start copying a (page-locked) numpy array into a TensorMemory on the upload stream; kernels queued after tf.cuda_wait_uploads() see it");
	m.def("cuda_wait_uploads", []() { Check(tfcuda_wait_uploads(), "cuda_wait_uploads"); },
	      "Order the backend stream behind every upload started so far (no host wait)");
	// a download in flight keeps its source tensor (and destination array) alive: references are parked here with the download's
	// ticket and dropped once the copy engine reports it complete
	// (heap-allocated and never destroyed: python objects must not be released after the interpreter has shut down)
	static auto& in_flight = *new std::deque<std::tuple<uint64_t, py::object, py::object>>();
	auto release_done = [](bool all) {
		uint64_t done = all ? UINT64_MAX : tfcuda_downloads_done();
		while (!in_flight.empty() && std::get<0>(in_flight.front()) <= done) in_flight.pop_front();
	};
	m.def("cuda_download_async", [release_done](py::object tensor, py::array arr) {
		RequireCuda("cuda_download_async");
		const PyTensorMemory& t = tensor.cast<const PyTensorMemory&>();
		py::buffer_info info = arr.request(true);
		(void)FormatOf(info);
		if (!(arr.flags() & py::array::c_style)) throw std::runtime_error("cuda_download_async: the destination array must be C-contiguous");
		if ((size_t)info.size != GetSize(t.tensor_)) throw std::runtime_error("cuda_download_async: element count mismatch");
		Check(tfcuda_memcpy_d2h_async(info.ptr, DevPtr(t), (size_t)info.size * 4), "cuda_download_async");
		in_flight.emplace_back(tfcuda_downloads_issued(), tensor, arr);
		release_done(false);
	}, "Start copying a TensorMemory into a (page-locked) numpy array on the download stream, after the work queued so far; "
	   "read it after tf.cuda_copy_sync().  The tensor is kept alive until the copy has finished");
	m.def("cuda_copy_sync", [release_done]() {
		Check(tfcuda_copy_sync(), "cuda_copy_sync");
		release_done(true);
	}, "Wait for the upload and download streams");
	m.def("cuda_graph_stats", []() {
		TFCudaGraphStats st{};
		tfcuda_graph_stats(&st);
		py::dict d;
		d["enabled"] = st.enabled != 0;
		d["replays"] = st.replays;
		d["exact_hits"] = st.exact_hits;
		d["patched"] = st.patched;
		d["instantiated"] = st.instantiated;
		d["eager_launches"] = st.eager_launches;
		d["host_ms"] = st.host_us / 1e3;
		return d;
	}, "Counters of the launch recorder (graph replay of program dispatch chains)");

	m.def("cuda_numpy", [](const PyTensorMemory& t) -> py::array {
		RequireCuda("cuda_numpy");
		std::vector<size_t> shape = t.Shape();
		py::array out;
		TFDataFormat f = t.GetFormat();
		if (f == TFTypeFloat32) out = py::array_t<float>(shape);
		else if (f == TFTypeInt32) out = py::array_t<int>(shape);
		else if (f == TFTypeUint32) out = py::array_t<uint>(shape);
		else throw std::runtime_error("cuda_numpy: bool tensors use .numpy");
		Check(tfcuda_memcpy_d2h(out.request().ptr, DevPtr(t), GetSize(t.tensor_) * 4), "cuda_numpy");
		return out;
	}, "numpy copy of a TensorMemory with one bulk copy");


	// ---- zero-copy export of device tensors: __cuda_array_interface__ v3 and DLPack on tf.TensorMemory ----
	{
		py::object cls = m.attr("TensorMemory");
		py::object property = py::module_::import("builtins").attr("property");
		cls.attr("__cuda_array_interface__") = property(py::cpp_function([](py::object self) {
			RequireCuda("__cuda_array_interface__");
			const PyTensorMemory& t = self.cast<const PyTensorMemory&>();
			py::dict d;
			std::vector<size_t> shape = t.Shape();
			py::tuple sh(shape.size());
			for (size_t i = 0; i < shape.size(); i++) sh[i] = shape[i];
			d["shape"] = sh;
			d["typestr"] = TypeStr(t.GetFormat());
			d["data"] = py::make_tuple((size_t)DevPtr(t), false);
			d["strides"] = py::none();
			d["version"] = 3;
			d["stream"] = (size_t)reinterpret_cast<uintptr_t>(tfcuda_stream());  // consumers order themselves behind the backend stream
			return d;
		}));
		cls.attr("__dlpack_device__") = py::cpp_function([](py::object) { return py::make_tuple(kDLCUDA, CurrentDevice()); }, py::is_method(cls));
		cls.attr("__dlpack__") = py::cpp_function([](py::object self, py::object stream) -> py::capsule {
			RequireCuda("__dlpack__");
			(void)stream;
			CudaFinish();  // the consumer may use any stream: hand over finished data
			const PyTensorMemory& t = self.cast<const PyTensorMemory&>();
			DLOwner* owner = new DLOwner();
			owner->tensor = self.ptr();
			Py_INCREF(owner->tensor);  // the exported view keeps the TensorMemory (and its device buffer) alive
			for (size_t d : t.Shape()) owner->shape.push_back((int64_t)d);
			DLTensor& dl = owner->managed.dl_tensor;
			dl.data = reinterpret_cast<void*>(DevPtr(t));
			dl.device = {kDLCUDA, CurrentDevice()};
			dl.ndim = (int32_t)owner->shape.size();
			TFDataFormat f = t.GetFormat();
			(void)TypeStr(f);
			dl.dtype = {(uint8_t)(f == TFTypeFloat32 ? 2 : (f == TFTypeInt32 ? 0 : 1)), 32, 1};
			dl.shape = owner->shape.data();
			dl.strides = nullptr;
			dl.byte_offset = 0;
			owner->managed.manager_ctx = owner;
			owner->managed.deleter = [](DLManagedTensor* mt) {
				DLOwner* o = static_cast<DLOwner*>(mt->manager_ctx);
				py::gil_scoped_acquire gil;
				Py_DECREF(o->tensor);
				delete o;
			};
			return py::capsule(&owner->managed, "dltensor", [](PyObject* cap) {
				// still named "dltensor": nobody consumed it, release here; a consumer renames it to "used_dltensor" and owns the deleter
				if (PyCapsule_IsValid(cap, "dltensor")) {
					auto* mt = static_cast<DLManagedTensor*>(PyCapsule_GetPointer(cap, "dltensor"));
					if (mt && mt->deleter) mt->deleter(mt);
				}
			});
		}, py::is_method(cls), py::arg("stream") = py::none());
	}
	m.def("cuda_from_device_array", [](py::object obj) {
		RequireCuda("cuda_from_device_array");
		if (!py::hasattr(obj, "__cuda_array_interface__")) throw std::runtime_error("cuda_from_device_array: the object has no __cuda_array_interface__");
		py::dict d = obj.attr("__cuda_array_interface__");
		std::string typestr = d["typestr"].cast<std::string>();
		TFDataFormat fmt;
		if (typestr == "<f4") fmt = TFTypeFloat32;
		else if (typestr == "<i4") fmt = TFTypeInt32;
		else if (typestr == "<u4") fmt = TFTypeUint32;
		else throw std::runtime_error("cuda_from_device_array: unsupported typestr " + typestr);
		if (d.contains("strides") && !d["strides"].is_none()) throw std::runtime_error("cuda_from_device_array: only C-contiguous arrays");
		std::vector<size_t> shape = d["shape"].cast<std::vector<size_t>>();
		size_t count = 1;
		for (size_t v : shape) count *= v;
		if (count == 0 || shape.empty()) throw std::runtime_error("cuda_from_device_array: empty arrays are not tensors");
		uint64_t src = d["data"].cast<py::tuple>()[0].cast<uint64_t>();
		TFTensor* t = global_memory_manager->AllocateTensor(shape, fmt);
		Check(tfcuda_sync(), "cuda_from_device_array");  // the producer synchronised before handing the pointer over (interface v2) or we cannot know its stream
		Check(tfcuda_memcpy_d2d(((TFCudaBuffer*)t->buffer)->GetNative(), src, count * 4), "cuda_from_device_array");
		return new PyTensorMemory(t);
	}, "TensorMemory from any object exposing __cuda_array_interface__ (torch / cupy / numba device arrays) with one device-to-device copy; "
	   "the producer must have finished writing (e.g. torch.cuda.synchronize())", py::return_value_policy::take_ownership);

	// ---- data-parallel exchange -------------------------------------------------------------------
	m.def("cuda_comm_unique_id", []() {
		uint8_t id[128];
		Check(tfcuda_comm_unique_id(id), "cuda_comm_unique_id");
		return py::bytes(reinterpret_cast<const char*>(id), 128);
	});
	m.def("cuda_comm_init", [](py::bytes id, int rank, int world) {
		RequireCuda("cuda_comm_init");
		std::string s = id;
		if (s.size() != 128) throw std::runtime_error("cuda_comm_init: unique id must be 128 bytes");
		Check(tfcuda_comm_init(reinterpret_cast<const uint8_t*>(s.data()), rank, world), "cuda_comm_init");
	});
	m.def("cuda_peer_export", []() {
		RequireCuda("cuda_peer_export");
		uint8_t h[64];
		Check(tfcuda_peer_export(h), "cuda_peer_export");
		return py::bytes(reinterpret_cast<const char*>(h), 64);
	}, "Allocate this rank's peer-memory exchange buffer and return its 64-byte CUDA IPC handle");
	m.def("cuda_peer_init", [](std::vector<py::bytes> handles, int rank, int world) {
		RequireCuda("cuda_peer_init");
		if ((int)handles.size() != world) throw std::runtime_error("cuda_peer_init: one handle per rank expected");
		std::string all;
		for (auto& h : handles) {
			std::string s = h;
			if (s.size() != 64) throw std::runtime_error("cuda_peer_init: IPC handles are 64 bytes");
			all += s;
		}
		Check(tfcuda_peer_init(reinterpret_cast<const uint8_t*>(all.data()), rank, world), "cuda_peer_init");
	}, "Map every peer's exchange buffer (handles in rank order)");
	m.def("cuda_peer_ready", []() { return tfcuda_peer_ready() != 0; });
	m.def("cuda_allreduce", [](PyTensorMemory& t, float scale, const std::string& method) {
		RequireCuda("cuda_allreduce");
		if (t.GetFormat() != TFTypeFloat32) throw std::runtime_error("cuda_allreduce: float32 tensors only");
		size_t count = GetSize(t.tensor_);
		bool peer = method == "peer" || (method == "auto" && tfcuda_peer_ready() && count <= tfcuda_peer_max_count());
		if (method != "peer" && method != "nccl" && method != "auto") throw std::runtime_error("cuda_allreduce: method must be auto, peer or nccl");
		if (peer) Check(tfcuda_peer_allreduce_sum_f32(DevPtr(t), count, scale), "cuda_allreduce (peer memory)");
		else Check(tfcuda_comm_allreduce_sum_f32(DevPtr(t), count, scale), "cuda_allreduce (NCCL)");
	}, py::arg("tensor"), py::arg("scale") = 1.0f, py::arg("method") = "auto",
	   "In-place sum-allreduce, then multiply by scale: one-shot kernel over NVLink peer memory for small tensors when the peer exchange "
	   "is initialised (tf.cuda_peer_init), NCCL otherwise");
	m.def("cuda_comm_destroy", []() { tfcuda_comm_destroy(); tfcuda_peer_destroy(); });

	// ---- library call inside a traced program: tf.sort.radix on this backend (see overlay/python/install.py) ----------
	m.def("cuda_library_active", []() {
		const char* v = getenv("TFCUDA_LIBRARY");
		return current_kernel_lang == CodeGenLang::CUDA && !(v && atoi(v) == 0);
	}, "True when algorithmic ops of traced programs are lowered to libtfcuda library calls (kernel language is CUDA)");
	m.def("cuda_library_sort", [](const PyTensor& keys, py::object values, int max_bits) -> py::object {
		const Tensor* v = values.is_none() ? nullptr : &values.cast<const PyTensor&>().Get();
		std::vector<Tensor*> out = CudaLibrarySort(&keys.Get(), v, max_bits);
		if (out.size() == 1) return py::cast(PT(*out[0]));
		return py::make_tuple(PT(*out[0]), PT(*out[1]));
	}, py::arg("keys"), py::arg("values") = py::none(), py::arg("max_bits") = 32,
	   "Traced stable LSD radix sort of a 1-D tensor (one libtfcuda library call in the compiled program)");

	// ---- library kernels on TensorMemory ------------------------------------------------------------
	m.def("cuda_radix_sort", [](const PyTensorMemory& keys, py::object values, int max_bits) -> py::object {
		RequireCuda("cuda_radix_sort");
		size_t n = GetSize(keys.tensor_);
		PyTensorMemory* keys_out = NewLike(keys, keys.GetFormat());
		PyTensorMemory temp({tfcuda_radix_sort_temp_words(n)}, TFTypeUint32);
		if (values.is_none()) {
			Check(tfcuda_radix_sort(DevPtr(keys), DevPtr(*keys_out), 0, 0, n, (int)keys.GetFormat().type, max_bits, DevPtr(temp)), "cuda_radix_sort");
			return py::cast(keys_out, py::return_value_policy::take_ownership);
		}
		const PyTensorMemory& vals = values.cast<const PyTensorMemory&>();
		if (GetSize(vals.tensor_) != n) throw std::runtime_error("cuda_radix_sort: keys and values differ in size");
		PyTensorMemory* vals_out = NewLike(vals, vals.GetFormat());
		Check(tfcuda_radix_sort(DevPtr(keys), DevPtr(*keys_out), DevPtr(vals), DevPtr(*vals_out), n, (int)keys.GetFormat().type, max_bits, DevPtr(temp)), "cuda_radix_sort");
		return py::make_tuple(py::cast(keys_out, py::return_value_policy::take_ownership), py::cast(vals_out, py::return_value_policy::take_ownership));
	}, py::arg("keys"), py::arg("values") = py::none(), py::arg("max_bits") = 32);

	m.def("cuda_reduce", [](const PyTensorMemory& in, int axis, const std::string& op) {
		RequireCuda("cuda_reduce");
		static const std::unordered_map<std::string, int> ops = {{"sum", TFCUDA_RED_SUM}, {"max", TFCUDA_RED_MAX}, {"min", TFCUDA_RED_MIN},
		    {"mean", TFCUDA_RED_MEAN}, {"norm", TFCUDA_RED_NORM}, {"prod", TFCUDA_RED_PROD}, {"any", TFCUDA_RED_ANY}, {"all", TFCUDA_RED_ALL}};
		auto it = ops.find(op);
		if (it == ops.end()) throw std::runtime_error("cuda_reduce: unknown op " + op);
		std::vector<size_t> shape = in.Shape();
		int dims = (int)shape.size();
		if (axis < 0) axis += dims;
		if (axis < 0 || axis >= dims) throw std::runtime_error("cuda_reduce: axis out of range");
		size_t outer = 1, inner = 1;
		for (int i = 0; i < axis; i++) outer *= shape[i];
		for (int i = axis + 1; i < dims; i++) inner *= shape[i];
		std::vector<size_t> out_shape;
		for (int i = 0; i < dims; i++) if (i != axis) out_shape.push_back(shape[i]);
		if (out_shape.empty()) out_shape.push_back(1);
		PyTensorMemory* out = new PyTensorMemory(global_memory_manager->AllocateTensor(out_shape, in.GetFormat()));
		Check(tfcuda_reduce(DevPtr(in), DevPtr(*out), outer, shape[axis], inner, it->second, (int)in.GetFormat().type), "cuda_reduce");
		return out;
	}, py::arg("tensor"), py::arg("axis") = -1, py::arg("op") = "sum", py::return_value_policy::take_ownership);

	m.def("cuda_prefix_sum", [](const PyTensorMemory& in, int axis) {
		RequireCuda("cuda_prefix_sum");
		std::vector<size_t> shape = in.Shape();
		int dims = (int)shape.size();
		if (axis < 0) axis += dims;
		if (axis < 0 || axis >= dims) throw std::runtime_error("cuda_prefix_sum: axis out of range");
		size_t outer = 1, inner = 1;
		for (int i = 0; i < axis; i++) outer *= shape[i];
		for (int i = axis + 1; i < dims; i++) inner *= shape[i];
		PyTensorMemory* out = NewLike(in, in.GetFormat());
		Check(tfcuda_prefix_sum(DevPtr(in), DevPtr(*out), outer, shape[axis], inner, (int)in.GetFormat().type), "cuda_prefix_sum");
		return out;
	}, py::arg("tensor"), py::arg("axis") = -1, py::return_value_policy::take_ownership);

	m.def("cuda_matmul", [](const PyTensorMemory& a, const PyTensorMemory& b, int mode) {
		RequireCuda("cuda_matmul");
		std::vector<size_t> sa = a.Shape(), sb = b.Shape();
		if (sa.size() < 2 || sb.size() != 2) throw std::runtime_error("cuda_matmul: expects A[..., M, K] @ B[K, N]");
		size_t k = sa.back(), mm = 1;
		for (size_t i = 0; i + 1 < sa.size(); i++) mm *= sa[i];
		if (sb[0] != k) throw std::runtime_error("cuda_matmul: inner dimensions differ");
		std::vector<size_t> so(sa.begin(), sa.end() - 1);
		so.push_back(sb[1]);
		PyTensorMemory* c = new PyTensorMemory(global_memory_manager->AllocateTensor(so, TFTypeFloat32));
		Check(tfcuda_matmul(DevPtr(a), DevPtr(b), DevPtr(*c), 1, mm, sb[1], k, mode), "cuda_matmul");
		return c;
	}, py::arg("a"), py::arg("b"), py::arg("mode") = 0, py::return_value_policy::take_ownership);

	m.def("cuda_scatter_add", [](PyTensorMemory& dst, const PyTensorMemory& index, const PyTensorMemory& src) {
		RequireCuda("cuda_scatter_add");
		size_t n = GetSize(src.tensor_);
		if (GetSize(index.tensor_) != n) throw std::runtime_error("cuda_scatter_add: index and src differ in size");
		Check(tfcuda_scatter_add(DevPtr(dst), DevPtr(index), DevPtr(src), n, GetSize(dst.tensor_), (int)dst.GetFormat().type), "cuda_scatter_add");
	});

	m.def("cuda_nbody_step", [](const PyTensorMemory& x, const PyTensorMemory& v, float dt, float eps) {
		RequireCuda("cuda_nbody_step");
		std::vector<size_t> s = x.Shape();
		if (s.size() != 2 || s[1] != 3) throw std::runtime_error("cuda_nbody_step: X must be [N,3]");
		PyTensorMemory* xn = NewLike(x, TFTypeFloat32);
		PyTensorMemory* vn = NewLike(v, TFTypeFloat32);
		Check(tfcuda_nbody_step(DevPtr(x), DevPtr(v), DevPtr(*xn), DevPtr(*vn), s[0], dt, eps), "cuda_nbody_step");
		return py::make_tuple(py::cast(xn, py::return_value_policy::take_ownership), py::cast(vn, py::return_value_policy::take_ownership));
	}, py::arg("x"), py::arg("v"), py::arg("dt") = 0.001f, py::arg("eps") = 1e-4f);
}

}  // namespace TensorFrost
