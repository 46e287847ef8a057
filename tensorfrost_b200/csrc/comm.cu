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

// ------------------------------------------------------------------------------------------------
// One-shot allreduce over NVLink peer memory.  The gradient exchange of the data-parallel NCA step is 7821 floats (31 KB):
// pure latency.  Every rank owns an exchange buffer [parity 2][rank][slot] + flags, exported with cudaIpcGetMemHandle and mapped
// by all peers.  ONE kernel per step, on the runtime stream right behind the gradient kernels:
//   push    every rank stores its vector into slot[parity][rank] of EVERY peer's buffer (NVSwitch: all peers at full rate),
//           __threadfence_system, then the last CTA to finish stores the step's epoch into flag[parity][rank] of every peer;
//   wait    spin until the own flags of all ranks show the epoch (the data of every peer has landed in local HBM);
//   reduce  sum the slots in RANK ORDER out of local memory (L2 loads: peer writes are coherent there), scale, write back.
// The rank-order sum makes the result bit-identical on all ranks (replicated optimizer state stays identical without a
// broadcast).  Slots alternate with the epoch's parity: a peer can only be one step ahead (it needs this rank's next flag to go
// further), so the slot it overwrites is never the one being read.  No NCCL call, no host synchronisation.
// ------------------------------------------------------------------------------------------------
constexpr int kPeerMaxRanks = 16;
constexpr size_t kPeerSlotFloats = 16384;             // 64 KB per rank and parity
constexpr size_t kPeerFlagStride = 32;                // uint32 words between flags (one 128-byte line each)
constexpr int kPeerBlocks = 16, kPeerThreads = 256;

struct PeerLayout {
	float* data[kPeerMaxRanks];      // base of every rank's exchange buffer as mapped in THIS process
	unsigned* flags[kPeerMaxRanks];
};

struct Peer {
	void* local = nullptr;           // this rank's buffer (cudaMalloc)
	void* mapped[kPeerMaxRanks] = {};
	PeerLayout layout{};
	unsigned* counter = nullptr;     // CTA arrival counter (local)
	int rank = -1, world = 0;
	unsigned epoch = 0;
};
Peer g_peer;

size_t peer_data_bytes() { return 2 * (size_t)kPeerMaxRanks * kPeerSlotFloats * sizeof(float); }
size_t peer_flag_bytes() { return 2 * (size_t)kPeerMaxRanks * kPeerFlagStride * sizeof(unsigned); }

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
	unsigned v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(float* buf, size_t count, float scale, PeerLayout L, unsigned* counter,
                                                                      int rank, int world, unsigned epoch) {
	const unsigned parity = epoch & 1u;
	const size_t slot = ((size_t)parity * kPeerMaxRanks + rank) * kPeerSlotFloats;
	const size_t n4 = count / 4;
	const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthreads = (size_t)gridDim.x * blockDim.x;
	// push
	for (size_t i = tid; i < n4; i += nthreads) {
		float4 v = reinterpret_cast<const float4*>(buf)[i];
		for (int p = 0; p < world; p++) reinterpret_cast<float4*>(L.data[p] + slot)[i] = v;
	}
	for (size_t i = n4 * 4 + tid; i < count; i += nthreads) {
		float v = buf[i];
		for (int p = 0; p < world; p++) L.data[p][slot + i] = v;
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned arrived = atomicAdd(counter, 1u);
		if (arrived == gridDim.x - 1) {
			*counter = 0;  // next launch starts from zero (launches are stream ordered)
			__threadfence_system();
			for (int p = 0; p < world; p++) st_release_sys(L.flags[p] + ((size_t)parity * kPeerMaxRanks + rank) * kPeerFlagStride, epoch);
		}
	}
	// wait for every rank's vector to land here
	if (threadIdx.x < world) {
		const unsigned* f = L.flags[rank] + ((size_t)parity * kPeerMaxRanks + threadIdx.x) * kPeerFlagStride;
		while (ld_acquire_sys(f) != epoch) {
		}
	}
	__syncthreads();
	// reduce in rank order from local memory (L2: __ldcg), scale, store
	const float* mine = L.data[rank] + (size_t)parity * kPeerMaxRanks * kPeerSlotFloats;
	for (size_t i = tid; i < n4; i += nthreads) {
		float4 acc = __ldcg(reinterpret_cast<const float4*>(mine) + i);
		for (int r = 1; r < world; r++) {
			float4 v = __ldcg(reinterpret_cast<const float4*>(mine + (size_t)r * kPeerSlotFloats) + i);
			acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
		}
		acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
		reinterpret_cast<float4*>(buf)[i] = acc;
	}
	for (size_t i = n4 * 4 + tid; i < count; i += nthreads) {
		float acc = __ldcg(mine + i);
		for (int r = 1; r < world; r++) acc += __ldcg(mine + (size_t)r * kPeerSlotFloats + i);
		buf[i] = acc * scale;
	}
}

}  // namespace

extern "C" {

int tfcuda_peer_export(uint8_t handle_out[64]) {
	if (!tfcuda::state().initialized) { tfcuda::set_error("tfcuda_peer_export: backend not initialised"); return 1; }
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
	if (!g_peer.local) {
		size_t bytes = peer_data_bytes() + peer_flag_bytes() + 256;
		TFCUDA_CHECK(cudaMalloc(&g_peer.local, bytes));
		TFCUDA_CHECK(cudaMemset(g_peer.local, 0, bytes));
		TFCUDA_CHECK(cudaDeviceSynchronize());
	}
	cudaIpcMemHandle_t h;
	TFCUDA_CHECK(cudaIpcGetMemHandle(&h, g_peer.local));
	memcpy(handle_out, &h, 64);
	return 0;
}

int tfcuda_peer_init(const uint8_t* handles, int rank, int world) {
	if (!g_peer.local) { tfcuda::set_error("tfcuda_peer_init: call tfcuda_peer_export first"); return 1; }
	if (world < 1 || world > kPeerMaxRanks || rank < 0 || rank >= world) { tfcuda::set_error("tfcuda_peer_init: rank/world out of range (at most 16 ranks)"); return 1; }
	if (g_peer.world) { tfcuda::set_error("tfcuda_peer_init: already initialised"); return 1; }
	for (int p = 0; p < world; p++) {
		void* base = g_peer.local;
		if (p != rank) {
			cudaIpcMemHandle_t h;
			memcpy(&h, handles + (size_t)p * 64, 64);
			TFCUDA_CHECK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
			g_peer.mapped[p] = base;
		}
		g_peer.layout.data[p] = reinterpret_cast<float*>(base);
		g_peer.layout.flags[p] = reinterpret_cast<unsigned*>(static_cast<char*>(base) + peer_data_bytes());
	}
	g_peer.counter = reinterpret_cast<unsigned*>(static_cast<char*>(g_peer.local) + peer_data_bytes() + peer_flag_bytes());
	g_peer.rank = rank;
	g_peer.world = world;
	g_peer.epoch = 0;
	return 0;
}

int tfcuda_peer_ready(void) { return g_peer.world > 0 ? 1 : 0; }

size_t tfcuda_peer_max_count(void) { return kPeerSlotFloats; }

int tfcuda_peer_allreduce_sum_f32(uint64_t buf, size_t count, float scale) {
	if (!g_peer.world) { tfcuda::set_error("tfcuda_peer_allreduce_sum_f32: no peer exchange (call tfcuda_peer_export / tfcuda_peer_init)"); return 1; }
	if (count > kPeerSlotFloats) { tfcuda::set_error("tfcuda_peer_allreduce_sum_f32: at most 16384 floats per exchange"); return 1; }
	if (buf % 16 != 0) { tfcuda::set_error("tfcuda_peer_allreduce_sum_f32: buffer must be 16-byte aligned"); return 1; }
	if (count == 0) return 0;
	cudaStream_t s = tfcuda::state().stream;
	g_peer.epoch++;
	if (g_peer.epoch == 0) g_peer.epoch = 2;  // 0 is the initial flag value; keep the parity sequence alternating
	tfcuda::ProfileScope prof("lib/peer_allreduce", 4.0 * (double)count * (g_peer.world + 1));
	peer_allreduce_kernel<<<kPeerBlocks, kPeerThreads, 0, s>>>(reinterpret_cast<float*>(buf), count, scale, g_peer.layout, g_peer.counter, g_peer.rank,
	                                                          g_peer.world, g_peer.epoch);
	return tfcuda::check_launch("peer_allreduce_kernel");
}

int tfcuda_peer_destroy(void) {
	if (!g_peer.local) return 0;
	cudaDeviceSynchronize();
	for (int p = 0; p < kPeerMaxRanks; p++)
		if (g_peer.mapped[p]) cudaIpcCloseMemHandle(g_peer.mapped[p]);
	cudaFree(g_peer.local);
	g_peer = Peer();
	return 0;
}

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
