// Data-parallel exchange for the NCA / neural-embedding training config: one NCCL communicator per
// process, a sum-allreduce of ONE flat fp32 gradient buffer on the runtime stream.
// The reference has no distributed code at all, so there is no counterpart to cite; the call sits
// between the program's gradient computation (tf.grad, tests/autograd_test.py:9-23 pattern) and its
// optimizer update (Python/TensorFrost/optimizers.py:121-147).
//
// NCCL is dlopen'ed (libnccl.so.2) so libtfcuda.so loads on boxes without it; when torch is already
// in the process the same soname resolves to torch's bundled copy.
#include <dlfcn.h>

#include "tfcuda_internal.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };

struct Nccl {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	ncclComm_t comm = nullptr;
	int world = 1;
};
Nccl g_nccl;

bool load_nccl() {
	if (g_nccl.handle) return true;
	const char* names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char* n : names) {
		g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		if (g_nccl.handle) break;
	}
	if (!g_nccl.handle) {
		tfcuda::set_error(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
		return false;
	}
	auto sym = [&](const char* s) { return dlsym(g_nccl.handle, s); };
	g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
	g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
	g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
	g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
	g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
		tfcuda::set_error("libnccl.so.2 lacks a required symbol");
		return false;
	}
	return true;
}

int nccl_fail(const char* what, ncclResult_t r) {
	tfcuda::set_error(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error") + " (" + std::to_string(r) + ")");
	return 1;
}

__global__ void scale_f32_kernel(float* p, size_t n, float s) {
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	if (i < n) p[i] *= s;
}

}  // namespace

extern "C" {

int tfcuda_comm_unique_id(uint8_t out[128]) {
	if (!load_nccl()) return 1;
	ncclUniqueId id;
	ncclResult_t r = g_nccl.GetUniqueId(&id);
	if (r) return nccl_fail("ncclGetUniqueId", r);
	memcpy(out, id.internal, 128);
	return 0;
}

int tfcuda_comm_init(const uint8_t unique_id[128], int rank, int world) {
	if (!tfcuda::state().initialized) { tfcuda::set_error("tfcuda_comm_init: backend not initialised"); return 1; }
	if (!load_nccl()) return 1;
	if (g_nccl.comm) { tfcuda::set_error("tfcuda_comm_init: communicator already exists"); return 1; }
	ncclUniqueId id;
	memcpy(id.internal, unique_id, 128);
	ncclResult_t r = g_nccl.CommInitRank(&g_nccl.comm, world, id, rank);
	if (r) return nccl_fail("ncclCommInitRank", r);
	g_nccl.world = world;
	return 0;
}

int tfcuda_comm_allreduce_sum_f32(uint64_t buf, size_t count, float scale) {
	if (!g_nccl.comm) { tfcuda::set_error("tfcuda_comm_allreduce_sum_f32: no communicator (call tfcuda_comm_init)"); return 1; }
	cudaStream_t s = tfcuda::state().stream;
	float* p = reinterpret_cast<float*>(buf);
	ncclResult_t r = g_nccl.AllReduce(p, p, count, ncclFloat32, ncclSum, g_nccl.comm, s);
	if (r) return nccl_fail("ncclAllReduce", r);
	tfcuda::state().launches++;
	if (scale != 1.0f && count) {
		scale_f32_kernel<<<(unsigned)((count + 255) / 256), 256, 0, s>>>(p, count, scale);
		return tfcuda::check_launch("scale_f32_kernel");
	}
	return 0;
}

int tfcuda_comm_destroy(void) {
	if (g_nccl.comm) {
		g_nccl.CommDestroy(g_nccl.comm);
		g_nccl.comm = nullptr;
	}
	return 0;
}

}  // extern "C"
