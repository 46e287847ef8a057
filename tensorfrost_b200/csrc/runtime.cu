// libtfcuda runtime: device/stream lifecycle, device buffers + pool, the TFRuntime callback table,
// NVRTC compilation of emitted kernels and their dispatch through the CUDA driver API.
//
// Reference counterparts (paths relative to the reference root):
//   InitializeBackend                       TensorFrost/Backend/Backend.cpp:10-73
//   Allocator..Region callbacks             TensorFrost/Backend/Backend.cpp:96-133
//   TensorMemoryManager pool                TensorFrost/Backend/TensorMemory.cpp:29-224
//   CpuMemoryManager / TFCPUBuffer          TensorFrost/Backend/Backends/CPU/Memory.h:18-59
//   CompileKernels / OpenGLKernelManager    TensorFrost/Backend/Backend.cpp:75-94, Backends/OpenGL/KernelManager.h:73-159
// Design differences: everything is ordered on ONE CUDA stream and only tf.read / readback synchronise;
// buffers come from the stream-ordered CUDA memory pool (cudaMallocAsync) so neither allocation nor
// release ever stalls the device; kernels of a program are compiled in parallel NVRTC chunks straight
// to sm_100a cubins and cached on disk.
#include <nvrtc.h>
#include <nvtx3/nvToolsExt.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include <unordered_map>
#include <vector>

#include "tfcuda_internal.h"

namespace tfcuda {

static State g_state;
static thread_local std::string g_error;
static std::string g_error_shared;
static std::mutex g_error_mutex;

State& state() { return g_state; }

void set_error(const std::string& msg) {
	g_error = msg;
	std::lock_guard<std::mutex> lock(g_error_mutex);
	g_error_shared = msg;
}

std::string cuda_err(cudaError_t e) {
	return std::string(cudaGetErrorName(e)) + " (" + cudaGetErrorString(e) + ")";
}

void require_init() {
	if (!g_state.initialized) {
		throw std::runtime_error("tfcuda: backend not initialised (tfcuda_init failed or was never called); there is no CPU fallback");
	}
}

// ------------------------------------------------------------------------------------------------
// Buffer pool used by the TFRuntime.alloc/dealloc callbacks (stand-alone use of the C-ABI).  Exact
// size-class free lists: programs have static shapes, so a released buffer is almost always requested
// again with the same size on the next step.  Capacity is rounded up to 64 words (256 B) so near-equal
// sizes share a class.  The device memory itself comes from cudaMallocAsync.
// ------------------------------------------------------------------------------------------------
struct Pool {
	std::unordered_map<size_t, std::vector<Buffer*>> free_lists;
	size_t allocated_words = 0;
	size_t unused_words = 0;
};
static Pool g_pool;

static size_t round_words(size_t words) { return (words + 63) & ~size_t(63); }

// Guard bands.  The reference compiler fuses a user load (Clamp mode) into the Unsafe load of its reduction lowering
// (Compiler/Implementations.cpp:293) and the clamp is lost: e.g. NCA's max_neighbor_alpha (examples/ML/NCA/nca.py:45-48) reads one
// image row before / after its tensor, on the reference's own C++ backend too.  What such a read returns is whatever lies next to
// the tensor, so results would depend on the allocation history of the process.  Every tensor buffer therefore sits between two
// zero-filled bands (zeroed once, when the buffer is created; emitted kernels never write out of range): reads that overshoot by
// less than the band see zeros - the value a fresh heap gives the reference - independent of what ran before.
static size_t guard_bytes() {
	static const size_t g = [] {
		const char* v = getenv("TFCUDA_GUARD_BYTES");
		size_t b = v ? (size_t)strtoull(v, nullptr, 0) : (size_t)16384;
		return (b + 255) & ~size_t(255);
	}();
	return g;
}

static void* g_scratch = nullptr;
static size_t g_scratch_bytes = 0;
static void flush_block_cache();

void* scratch(size_t bytes) {
	if (bytes <= g_scratch_bytes) return g_scratch;
	if (g_scratch) cudaFreeAsync(g_scratch, g_state.stream);
	g_scratch = nullptr;
	g_scratch_bytes = 0;
	size_t want = (bytes + (bytes >> 3) + 0xfffff) & ~size_t(0xfffff);
	void* p = nullptr;
	cudaError_t e = cudaMallocAsync(&p, want, g_state.stream);
	if (e != cudaSuccess) {
		// parked tensor blocks / the driver pool may hold what the scratch needs: release them and ask for the exact size
		(void)cudaGetLastError();
		flush_block_cache();
		cudaStreamSynchronize(g_state.stream);
		cudaMemPool_t mp;
		if (cudaDeviceGetDefaultMemPool(&mp, g_state.device) == cudaSuccess) cudaMemPoolTrimTo(mp, 0);
		(void)cudaGetLastError();
		want = bytes;
		e = cudaMallocAsync(&p, want, g_state.stream);
	}
	if (e != cudaSuccess) {
		set_error("tfcuda: scratch allocation of " + std::to_string(bytes) + " bytes failed: " + cuda_err(e));
		(void)cudaGetLastError();
		return nullptr;
	}
	g_scratch = p;
	g_scratch_bytes = want;
	return p;
}

// Block cache under create_buffer / destroy_buffer.  The reference's TensorMemoryManager deletes a buffer that stayed unused for
// more than ~128 ALLOCATIONS (Backend/TensorMemory.cpp:145-160, one tick per tf.allocate) and re-creates it on the next step; a
// training step with 760 allocations therefore keeps creating and deleting multi-GB buffers.  Each cudaMallocAsync / cudaFreeAsync
// of that size costs the driver a virtual-memory operation (measured: ~0.6 ms a pair at 2-4 GB, and they serialise between the
// processes of a multi-GPU job), so deleted blocks are parked here by exact size and handed back without a driver call.  The cache
// is flushed when it exceeds TFCUDA_BLOCK_CACHE_MB (default: a quarter of device memory) or when an allocation fails.
struct BlockCache {
	std::unordered_map<size_t, std::vector<void*>> blocks;  // total bytes -> parked blocks (guard bands still zero)
	size_t bytes = 0;
	size_t limit = 0;
	uint64_t driver_calls = 0, hits = 0;
};
static BlockCache g_blocks;

static void flush_block_cache() {
	for (auto& kv : g_blocks.blocks)
		for (void* p : kv.second) {
			cudaFreeAsync(p, g_state.stream);
			g_blocks.driver_calls++;
		}
	g_blocks.blocks.clear();
	g_blocks.bytes = 0;
}

static Buffer* create_buffer(size_t words) {
	require_init();
	if (words == 0) throw std::invalid_argument("tfcuda: trying to allocate a buffer with size 0");
	void* p = nullptr;
	const size_t guard = guard_bytes();
	const size_t payload = (words * sizeof(uint32_t) + 255) & ~size_t(255);
	const size_t total = payload + 2 * guard;
	auto hit = g_blocks.blocks.find(total);
	if (hit != g_blocks.blocks.end() && !hit->second.empty()) {
		p = hit->second.back();
		hit->second.pop_back();
		g_blocks.bytes -= total;
		g_blocks.hits++;
		// the previous tenant may have been up to 63 words longer (same 256-byte class): re-zero the slack behind this tensor
		if (payload != words * sizeof(uint32_t))
			cudaMemsetAsync(static_cast<char*>(p) + guard + words * sizeof(uint32_t), 0, payload - words * sizeof(uint32_t), g_state.stream);
	} else {
		cudaError_t e = cudaMallocAsync(&p, total, g_state.stream);
		g_blocks.driver_calls++;
		if (e != cudaSuccess) {
			// parked blocks / the driver pool may be holding memory another size could use: release and retry once
			flush_block_cache();
			cudaStreamSynchronize(g_state.stream);
			cudaMemPool_t mp;
			if (cudaDeviceGetDefaultMemPool(&mp, g_state.device) == cudaSuccess) cudaMemPoolTrimTo(mp, 0);
			(void)cudaGetLastError();
			e = cudaMallocAsync(&p, total, g_state.stream);
			g_blocks.driver_calls++;
		}
		if (e != cudaSuccess) {
			std::string m = "tfcuda: device allocation of " + std::to_string(words * 4) + " bytes failed: " + cuda_err(e);
			set_error(m);
			throw std::runtime_error(m);
		}
		if (guard) {
			cudaMemsetAsync(p, 0, guard, g_state.stream);
			cudaMemsetAsync(static_cast<char*>(p) + guard + words * sizeof(uint32_t), 0, total - guard - words * sizeof(uint32_t), g_state.stream);
		}
	}
	Buffer* b = new Buffer();
	b->base.size = words;
	b->dptr = reinterpret_cast<uint64_t>(p) + guard;
	g_pool.allocated_words += words;
	return b;
}

static void destroy_buffer(Buffer* b) {
	if (!b) return;
	if (b->dptr && g_state.initialized) {
		const size_t guard = guard_bytes();
		const size_t total = ((b->base.size * sizeof(uint32_t) + 255) & ~size_t(255)) + 2 * guard;
		void* p = reinterpret_cast<void*>(b->dptr - guard);
		if (g_blocks.limit == 0) {
			const char* v = getenv("TFCUDA_BLOCK_CACHE_MB");
			size_t free_b = 0, total_b = 0;
			cudaMemGetInfo(&free_b, &total_b);
			g_blocks.limit = v ? (size_t)strtoull(v, nullptr, 0) << 20 : total_b / 4;
			if (g_blocks.limit == 0) g_blocks.limit = 1;  // TFCUDA_BLOCK_CACHE_MB=0: no parking
		}
		if (total <= g_blocks.limit) {
			if (g_blocks.bytes + total > g_blocks.limit) flush_block_cache();
			g_blocks.blocks[total].push_back(p);
			g_blocks.bytes += total;
		} else {
			cudaFreeAsync(p, g_state.stream);
			g_blocks.driver_calls++;
		}
	}
	g_pool.allocated_words -= b->base.size;
	delete b;
}

// ------------------------------------------------------------------------------------------------
// TFRuntime callbacks
// ------------------------------------------------------------------------------------------------
static TFTensor rt_alloc(const char* name, const size_t* shape, size_t dim, TFDataFormat fmt, void*) {
	size_t words = 1;
	for (size_t i = 0; i < dim; i++) words *= shape[i];
	if (words == 0) throw std::invalid_argument(std::string("tfcuda: tensor ") + (name ? name : "?") + " has size 0");
	size_t cls = round_words(words);
	Buffer* b = nullptr;
	auto it = g_pool.free_lists.find(cls);
	if (it != g_pool.free_lists.end() && !it->second.empty()) {
		b = it->second.back();
		it->second.pop_back();
		g_pool.unused_words -= b->base.size;
	} else {
		b = create_buffer(cls);
	}
	b->base.used_size = words;
	b->base.time_since_used = 0;
	b->base.read_only = false;
	b->base.up_to_date = false;
	b->base.name = name;
	size_t* shape_copy = new size_t[dim ? dim : 1];
	for (size_t i = 0; i < dim; i++) shape_copy[i] = shape[i];
	TFTensor t;
	t.buffer = &b->base;
	t.format = fmt;
	t.dim = dim;
	t.shape = shape_copy;
	return t;
}

static void rt_dealloc(TFTensor t, void*) {
	if (!t.buffer) return;
	Buffer* b = reinterpret_cast<Buffer*>(t.buffer);
	b->base.used_size = 0;
	b->base.name = "none";
	g_pool.free_lists[b->base.size].push_back(b);
	g_pool.unused_words += b->base.size;
}

static uint32_t rt_readback(TFTensor t, size_t index, void*) {
	require_init();
	if (!t.buffer || index >= t.buffer->size) throw std::out_of_range("tfcuda: tf.read index out of range");
	TFCUDA_THROW(cudaMemcpyAsync(g_state.pinned_word, reinterpret_cast<const void*>(dptr_of(t.buffer) + index * 4), 4,
	                             cudaMemcpyDeviceToHost, g_state.stream));
	TFCUDA_THROW(cudaStreamSynchronize(g_state.stream));
	return *g_state.pinned_word;
}

static void rt_writeback(TFTensor t, size_t index, uint32_t value, void*) {
	require_init();
	if (!t.buffer || index >= t.buffer->size) throw std::out_of_range("tfcuda: tf.write index out of range");
	// a 32-bit fill is stream ordered and needs no host staging
	int rc = tfcuda_memset32(dptr_of(t.buffer) + index * 4, value, 1);
	if (rc) throw std::runtime_error(std::string("tfcuda: tf.write failed: ") + tfcuda_last_error());
}

static void rt_dispatch(TFDispatchInfo info, void*) {
	if (tfcuda_dispatch(&info)) throw std::runtime_error(std::string("tfcuda: dispatch of kernel ") + std::to_string(info.kernel_id) + " failed: " + tfcuda_last_error());
}

static void rt_region(const char* name, bool begin, void*) {
	if (begin) nvtxRangePushA(name ? name : "region");
	else nvtxRangePop();
}

// ------------------------------------------------------------------------------------------------
// Kernel registry
// ------------------------------------------------------------------------------------------------
struct KernelEntry {
	CUfunction fn = nullptr;
	unsigned group[3] = {1, 1, 1};
	unsigned n_mem = 0;
	unsigned n_var = 0;
	unsigned library_op = 0;
	std::string entry;
};
static std::vector<KernelEntry> g_kernels;
static std::vector<CUmodule> g_modules;

static uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
	for (unsigned char c : s) {
		h ^= c;
		h *= 1099511628211ull;
	}
	return h;
}

static std::string cache_dir() {
	const char* env = getenv("TFCUDA_CACHE_DIR");
	std::string dir = env ? env : ("/tmp/tfcuda_cache_" + std::to_string((long)getuid()));
	mkdir(dir.c_str(), 0700);
	return dir;
}

static bool read_file(const std::string& path, std::string& out) {
	std::ifstream f(path, std::ios::binary);
	if (!f) return false;
	std::stringstream ss;
	ss << f.rdbuf();
	out = ss.str();
	return !out.empty();
}

static void write_file_atomic(const std::string& path, const std::string& data) {
	std::string tmp = path + ".tmp" + std::to_string((long)getpid());
	{
		std::ofstream f(tmp, std::ios::binary);
		if (!f) return;
		f.write(data.data(), (std::streamsize)data.size());
	}
	rename(tmp.c_str(), path.c_str());
}

static const char kPrelude[] =
#include "prelude_embed.inc"
    ;

struct Chunk {
	std::string source;
	std::string cubin;
	std::string log;
	bool ok = false;
	bool from_cache = false;
};

static void compile_chunk(Chunk& c, const std::vector<std::string>& opts, bool use_cache) {
	std::string key_src = c.source;
	for (auto& o : opts) key_src += "\x01" + o;
	int maj = 0, min = 0;
	nvrtcVersion(&maj, &min);
	key_src += "\x01nvrtc" + std::to_string(maj) + "." + std::to_string(min);
	char name[64];
	snprintf(name, sizeof(name), "%016llx.cubin", (unsigned long long)fnv1a(key_src));
	std::string path = use_cache ? cache_dir() + "/" + name : std::string();
	if (use_cache && read_file(path, c.cubin)) {
		c.ok = true;
		c.from_cache = true;
		return;
	}
	nvrtcProgram prog = nullptr;
	nvrtcResult r = nvrtcCreateProgram(&prog, c.source.c_str(), "tf_kernels.cu", 0, nullptr, nullptr);
	if (r != NVRTC_SUCCESS) {
		c.log = std::string("nvrtcCreateProgram: ") + nvrtcGetErrorString(r);
		return;
	}
	std::vector<const char*> copts;
	for (auto& o : opts) copts.push_back(o.c_str());
	r = nvrtcCompileProgram(prog, (int)copts.size(), copts.data());
	size_t log_size = 0;
	nvrtcGetProgramLogSize(prog, &log_size);
	if (log_size > 1) {
		c.log.resize(log_size);
		nvrtcGetProgramLog(prog, c.log.data());
	}
	if (r != NVRTC_SUCCESS) {
		c.log = std::string("nvrtcCompileProgram: ") + nvrtcGetErrorString(r) + "\n" + c.log;
		nvrtcDestroyProgram(&prog);
		return;
	}
	size_t size = 0;
	nvrtcGetCUBINSize(prog, &size);
	c.cubin.resize(size);
	nvrtcGetCUBIN(prog, c.cubin.data());
	nvrtcDestroyProgram(&prog);
	c.ok = size > 0;
	if (!c.ok) c.log += "\nempty cubin";
	if (c.ok && use_cache) write_file_atomic(path, c.cubin);
}

// ------------------------------------------------------------------------------------------------
// Profiling
// ------------------------------------------------------------------------------------------------
struct ProfileSample { cudaEvent_t a, b; };
struct ProfileEntry {
	std::vector<ProfileSample> pending;
	uint64_t launches = 0;
	double total_ms = 0.0;
	double bytes = 0.0;
};
static bool g_profile_on = false;
static std::unordered_map<std::string, ProfileEntry> g_profile;
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t take_event() {
	if (!g_event_pool.empty()) {
		cudaEvent_t e = g_event_pool.back();
		g_event_pool.pop_back();
		return e;
	}
	cudaEvent_t e;
	cudaEventCreate(&e);
	return e;
}

void profile_begin(const char* name) {
	if (!g_profile_on) return;
	ProfileEntry& pe = g_profile[name];
	ProfileSample smp{take_event(), take_event()};
	cudaEventRecord(smp.a, g_state.stream);
	pe.pending.push_back(smp);
}

void profile_end(const char* name, double bytes) {
	if (!g_profile_on) return;
	ProfileEntry& pe = g_profile[name];
	if (pe.pending.empty()) return;
	cudaEventRecord(pe.pending.back().b, g_state.stream);
	pe.launches++;
	pe.bytes += bytes;
}

static void profile_resolve() {
	cudaStreamSynchronize(g_state.stream);
	for (auto& kv : g_profile) {
		for (ProfileSample& smp : kv.second.pending) {
			float ms = 0;
			if (cudaEventElapsedTime(&ms, smp.a, smp.b) == cudaSuccess) kv.second.total_ms += ms;
			g_event_pool.push_back(smp.a);
			g_event_pool.push_back(smp.b);
		}
		kv.second.pending.clear();
	}
	(void)cudaGetLastError();
}

static std::string drv_err(CUresult r) {
	const char* s = nullptr;
	if (g_state.drv.GetErrorString) g_state.drv.GetErrorString(r, &s);
	return s ? s : ("CUresult " + std::to_string((int)r));
}

template <typename T>
static bool load_entry(const char* sym, T& fn) {
	void* p = nullptr;
	cudaDriverEntryPointQueryResult q;
	cudaError_t e = cudaGetDriverEntryPoint(sym, &p, cudaEnableDefault, &q);
	if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
		set_error(std::string("cannot resolve driver entry point ") + sym);
		return false;
	}
	fn = reinterpret_cast<T>(p);
	return true;
}

}  // namespace tfcuda

using namespace tfcuda;

// ================================================================================================
// C-ABI
// ================================================================================================
extern "C" {

const char* tfcuda_last_error(void) {
	if (!g_error.empty()) return g_error.c_str();
	std::lock_guard<std::mutex> lock(g_error_mutex);
	g_error = g_error_shared;
	return g_error.c_str();
}

const char* tfcuda_prelude(void) { return kPrelude; }

int tfcuda_is_initialized(void) { return g_state.initialized ? 1 : 0; }

int tfcuda_init(int device) {
	if (g_state.initialized) {
		if (device >= 0 && device != g_state.device) {
			set_error("tfcuda_init: already initialised on device " + std::to_string(g_state.device) + " (one device per process)");
			return 1;
		}
		return 0;
	}
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		(void)cudaGetLastError();
		set_error("tfcuda_init: no CUDA device available (" + (e != cudaSuccess ? cuda_err(e) : std::string("device count 0")) + "); this backend has no CPU fallback");
		return 1;
	}
	if (device < 0) {
		const char* lr = getenv("LOCAL_RANK");
		device = lr ? atoi(lr) % count : 0;
	}
	if (device >= count) {
		set_error("tfcuda_init: device " + std::to_string(device) + " out of range, " + std::to_string(count) + " visible");
		return 1;
	}
	TFCUDA_CHECK(cudaSetDevice(device));
	TFCUDA_CHECK(cudaFree(0));  // force primary-context creation
	cudaDeviceProp prop;
	TFCUDA_CHECK(cudaGetDeviceProperties(&prop, device));
	g_state.device = device;
	g_state.sm_count = prop.multiProcessorCount;
	g_state.device_name = prop.name;
	if (prop.major < 10 && !getenv("TFCUDA_ALLOW_ANY_ARCH")) {
		set_error("tfcuda_init: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + "; this backend is built for sm_100a (B200) only");
		return 1;
	}
	DriverApi& d = g_state.drv;
	if (!load_entry("cuModuleLoadData", d.ModuleLoadData) || !load_entry("cuModuleUnload", d.ModuleUnload) ||
	    !load_entry("cuModuleGetFunction", d.ModuleGetFunction) || !load_entry("cuLaunchKernel", d.LaunchKernel) ||
	    !load_entry("cuGetErrorString", d.GetErrorString) || !load_entry("cuFuncGetAttribute", d.FuncGetAttribute)) {
		return 1;
	}
	if (getenv("TFCUDA_PDL") != nullptr && atoi(getenv("TFCUDA_PDL")) != 0) {
		// optional entry point of the experimental programmatic-dependent-launch path; without it launches stay ordinary
		void* p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuLaunchKernelEx", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && p)
			d.LaunchKernelEx = reinterpret_cast<decltype(d.LaunchKernelEx)>(p);
		(void)cudaGetLastError();
	}
	TFCUDA_CHECK(cudaStreamCreateWithFlags(&g_state.stream, cudaStreamNonBlocking));
	TFCUDA_CHECK(cudaMallocHost(&g_state.pinned_word, 64));
	TFCUDA_CHECK(cudaEventCreate(&g_state.ev_begin));
	TFCUDA_CHECK(cudaEventCreate(&g_state.ev_end));
	// keep freed blocks in the stream-ordered pool instead of returning them to the OS at every sync
	cudaMemPool_t mp;
	TFCUDA_CHECK(cudaDeviceGetDefaultMemPool(&mp, device));
	uint64_t threshold = UINT64_MAX;
	TFCUDA_CHECK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &threshold));
	g_state.initialized = true;
	return 0;
}

int tfcuda_shutdown(void) {
	if (!g_state.initialized) return 0;
	cudaStreamSynchronize(g_state.stream);
	for (auto& kv : g_pool.free_lists)
		for (Buffer* b : kv.second) destroy_buffer(b);
	g_pool.free_lists.clear();
	g_pool.unused_words = 0;
	flush_block_cache();
	if (g_scratch) cudaFreeAsync(g_scratch, g_state.stream);
	g_scratch = nullptr;
	g_scratch_bytes = 0;
	for (CUmodule m : g_modules) g_state.drv.ModuleUnload(m);
	g_modules.clear();
	g_kernels.clear();
	cudaEventDestroy(g_state.ev_begin);
	cudaEventDestroy(g_state.ev_end);
	cudaFreeHost(g_state.pinned_word);
	cudaStreamDestroy(g_state.stream);
	g_state = State();
	return 0;
}

int tfcuda_device_sm_count(void) { return g_state.sm_count; }
const char* tfcuda_device_name(void) { return g_state.device_name.c_str(); }
void* tfcuda_stream(void) { return g_state.stream; }

int tfcuda_sync(void) {
	if (!g_state.initialized) {
		set_error("tfcuda_sync: not initialised");
		return 1;
	}
	TFCUDA_CHECK(cudaStreamSynchronize(g_state.stream));
	return 0;
}

TFRuntime tfcuda_runtime(void) {
	TFRuntime rt;
	rt.alloc = rt_alloc;
	rt.dealloc = rt_dealloc;
	rt.readback = rt_readback;
	rt.writeback = rt_writeback;
	rt.dispatch = rt_dispatch;
	rt.region = rt_region;
	rt.custom_data = nullptr;
	return rt;
}

// ---- buffers ------------------------------------------------------------------------------------
TFBuffer* tfcuda_buffer_create(size_t words) {
	try {
		return &create_buffer(words)->base;
	} catch (const std::exception& e) {
		set_error(e.what());
		return nullptr;
	}
}

void tfcuda_buffer_destroy(TFBuffer* buffer) { destroy_buffer(reinterpret_cast<Buffer*>(buffer)); }

uint64_t tfcuda_buffer_device_ptr(const TFBuffer* buffer) { return buffer ? dptr_of(buffer) : 0; }

int tfcuda_buffer_write(TFBuffer* buffer, size_t word_offset, const uint32_t* src, size_t words) {
	if (!buffer || word_offset + words > buffer->size) {
		set_error("tfcuda_buffer_write: range exceeds buffer");
		return 1;
	}
	return tfcuda_memcpy_h2d(dptr_of(buffer) + word_offset * 4, src, words * 4);
}

int tfcuda_buffer_read(const TFBuffer* buffer, size_t word_offset, uint32_t* dst, size_t words) {
	if (!buffer || word_offset + words > buffer->size) {
		set_error("tfcuda_buffer_read: range exceeds buffer");
		return 1;
	}
	return tfcuda_memcpy_d2h(dst, dptr_of(buffer) + word_offset * 4, words * 4);
}

int tfcuda_memcpy_h2d(uint64_t dst, const void* src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	TFCUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, g_state.stream));
	// pageable sources are consumed before the call returns; pinned ones are not, so order the host too
	TFCUDA_CHECK(cudaStreamSynchronize(g_state.stream));
	return 0;
}

int tfcuda_memcpy_d2h(void* dst, uint64_t src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	TFCUDA_CHECK(cudaMemcpyAsync(dst, reinterpret_cast<const void*>(src), bytes, cudaMemcpyDeviceToHost, g_state.stream));
	TFCUDA_CHECK(cudaStreamSynchronize(g_state.stream));
	return 0;
}

int tfcuda_memcpy_d2d(uint64_t dst, uint64_t src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	TFCUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(dst), reinterpret_cast<const void*>(src), bytes, cudaMemcpyDeviceToDevice, g_state.stream));
	return 0;
}

uint64_t tfcuda_malloc(size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 0; }
	void* p = nullptr;
	cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 4, g_state.stream);
	if (e != cudaSuccess) {
		set_error("tfcuda_malloc(" + std::to_string(bytes) + "): " + cuda_err(e));
		return 0;
	}
	return reinterpret_cast<uint64_t>(p);
}

int tfcuda_free(uint64_t ptr) {
	if (!ptr) return 0;
	TFCUDA_CHECK(cudaFreeAsync(reinterpret_cast<void*>(ptr), g_state.stream));
	return 0;
}

size_t tfcuda_pool_allocated_words(void) { return g_pool.allocated_words; }
size_t tfcuda_pool_unused_words(void) { return g_pool.unused_words; }
uint64_t tfcuda_pool_driver_calls(void) { return g_blocks.driver_calls; }

// ---- kernels ------------------------------------------------------------------------------------
static std::vector<std::string> nvrtc_options(const char* options) {
	std::vector<std::string> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-default-device", "-lineinfo"};
	if (options) {
		std::istringstream ss(options);
		std::string tok;
		while (ss >> tok) opts.push_back(tok);
	}
	return opts;
}

int tfcuda_nvrtc_check(const char* source, const char* options) {
	Chunk c;
	c.source = std::string(kPrelude) + "\n" + (source ? source : "");
	compile_chunk(c, nvrtc_options(options), false);
	if (!c.ok) {
		set_error("NVRTC: " + c.log);
		return 2;
	}
	return 0;
}

int tfcuda_compile_kernels(const TFCudaKernelSource* kernels, size_t count, const char* options) {
	if (!g_state.initialized) { set_error("tfcuda_compile_kernels: not initialised"); return 1; }
	if (count == 0) return 0;

	std::vector<std::string> opts = nvrtc_options(options);
	bool use_cache = getenv("TFCUDA_NO_CACHE") == nullptr;

	// chunk the emitted kernels; each chunk is one NVRTC translation unit = prelude + kernels
	const size_t kPerChunk = 12;
	std::vector<size_t> emitted;
	for (size_t i = 0; i < count; i++)
		if (kernels[i].library_op == 0) emitted.push_back(i);
	std::vector<Chunk> chunks((emitted.size() + kPerChunk - 1) / kPerChunk);
	for (size_t c = 0; c < chunks.size(); c++) {
		std::string& src = chunks[c].source;
		src = kPrelude;
		for (size_t j = c * kPerChunk; j < std::min(emitted.size(), (c + 1) * kPerChunk); j++) {
			src += "\n";
			src += kernels[emitted[j]].source;
		}
	}
	unsigned hw = std::thread::hardware_concurrency();
	size_t n_threads = std::min<size_t>(chunks.size(), hw ? hw : 4);
	std::atomic<size_t> next{0};
	auto worker = [&]() {
		for (;;) {
			size_t i = next.fetch_add(1);
			if (i >= chunks.size()) break;
			compile_chunk(chunks[i], opts, use_cache);
		}
	};
	if (n_threads <= 1) {
		worker();
	} else {
		std::vector<std::thread> threads;
		for (size_t t = 0; t < n_threads; t++) threads.emplace_back(worker);
		for (auto& t : threads) t.join();
	}

	for (size_t c = 0; c < chunks.size(); c++) {
		if (!chunks[c].ok) {
			set_error("NVRTC failed for kernel chunk " + std::to_string(c) + ":\n" + chunks[c].log + "\n---- source ----\n" + chunks[c].source.substr(strlen(kPrelude)));
			return 2;
		}
	}

	std::vector<CUmodule> mods(chunks.size(), nullptr);
	for (size_t c = 0; c < chunks.size(); c++) {
		CUresult r = g_state.drv.ModuleLoadData(&mods[c], chunks[c].cubin.data());
		if (r != CUDA_SUCCESS) {
			set_error("cuModuleLoadData: " + drv_err(r));
			return 3;
		}
		g_modules.push_back(mods[c]);
	}
	for (size_t i = 0; i < count; i++) {
		const TFCudaKernelSource& k = kernels[i];
		if (k.kernel_id >= g_kernels.size()) g_kernels.resize(k.kernel_id + 1);
		KernelEntry& e = g_kernels[k.kernel_id];
		e = KernelEntry();
		for (int d = 0; d < 3; d++) e.group[d] = k.group[d] ? k.group[d] : 1;
		e.n_mem = k.n_mem;
		e.n_var = k.n_var;
		e.library_op = k.library_op;
		e.entry = k.entry ? k.entry : "";
	}
	for (size_t j = 0; j < emitted.size(); j++) {
		const TFCudaKernelSource& k = kernels[emitted[j]];
		KernelEntry& e = g_kernels[k.kernel_id];
		CUresult r = g_state.drv.ModuleGetFunction(&e.fn, mods[j / kPerChunk], k.entry);
		if (r != CUDA_SUCCESS) {
			set_error(std::string("cuModuleGetFunction(") + k.entry + "): " + drv_err(r));
			return 4;
		}
	}
	return 0;
}

int tfcuda_launch(size_t kernel_id, const uint64_t* mem, size_t n_mem, const uint32_t* vars, size_t n_var, size_t work_group_count) {
	if (!g_state.initialized) { set_error("tfcuda_launch: not initialised"); return 1; }
	if (kernel_id >= g_kernels.size() || (!g_kernels[kernel_id].fn && !g_kernels[kernel_id].library_op)) {
		set_error("tfcuda_launch: kernel " + std::to_string(kernel_id) + " was never compiled");
		return 1;
	}
	KernelEntry& e = g_kernels[kernel_id];
	if (e.library_op) {
		set_error("tfcuda_launch: kernel " + std::to_string(kernel_id) + " is a library call; it is dispatched by the backend glue (tfcuda_matmul / tfcuda_reduce / ...), not launched");
		return 1;
	}
	if (n_mem != e.n_mem || n_var != e.n_var) {
		set_error("tfcuda_launch: kernel " + std::to_string(kernel_id) + " expects " + std::to_string(e.n_mem) + " buffers / " + std::to_string(e.n_var) +
		          " variables, got " + std::to_string(n_mem) + " / " + std::to_string(n_var));
		return 1;
	}
	if (work_group_count == 0) return 0;
	// argument block = { uint* mem[n_mem]; uint var[n_var]; } passed by value as the single kernel parameter
	alignas(8) unsigned char block[4096];
	size_t bytes = n_mem * 8 + n_var * 4;
	if (bytes > sizeof(block)) {
		set_error("tfcuda_launch: argument block too large");
		return 1;
	}
	memcpy(block, mem, n_mem * 8);
	memcpy(block + n_mem * 8, vars, n_var * 4);
	uint32_t* offset_word = n_var ? reinterpret_cast<uint32_t*>(block + n_mem * 8 + (n_var - 1) * 4) : nullptr;
	void* params[1] = {block};
	// grid.x is limited to 2^31-1: larger dispatches are split with the _kernel_block_offset word
	const size_t kMaxGrid = 0x7fffffffull;
	size_t done = 0;
	ProfileScope prof(e.entry.c_str());
	static const bool pdl_env = getenv("TFCUDA_PDL") != nullptr && atoi(getenv("TFCUDA_PDL")) != 0;
	const bool pdl = pdl_env && g_state.drv.LaunchKernelEx != nullptr;
	while (done < work_group_count) {
		size_t now = std::min(kMaxGrid, work_group_count - done);
		if (offset_word) *offset_word = vars[n_var - 1] + (uint32_t)done;
		CUresult r;
		if (pdl) {
			// experimental: the kernel begins with griddepcontrol.wait (emitted under the same switch), so it may be scheduled while its
			// predecessor drains; without the switch this branch is never taken
			CUlaunchAttribute attr;
			memset(&attr, 0, sizeof(attr));
			attr.id = CU_LAUNCH_ATTRIBUTE_PROGRAMMATIC_STREAM_SERIALIZATION;
			attr.value.programmaticStreamSerializationAllowed = 1;
			CUlaunchConfig cfg;
			memset(&cfg, 0, sizeof(cfg));
			cfg.gridDimX = (unsigned)now; cfg.gridDimY = 1; cfg.gridDimZ = 1;
			cfg.blockDimX = e.group[0]; cfg.blockDimY = e.group[1]; cfg.blockDimZ = e.group[2];
			cfg.sharedMemBytes = 0;
			cfg.hStream = (CUstream)g_state.stream;
			cfg.attrs = &attr;
			cfg.numAttrs = 1;
			r = g_state.drv.LaunchKernelEx(&cfg, e.fn, params, nullptr);
		} else {
			r = g_state.drv.LaunchKernel(e.fn, (unsigned)now, 1, 1, e.group[0], e.group[1], e.group[2], 0, (CUstream)g_state.stream, params, nullptr);
		}
		if (r != CUDA_SUCCESS) {
			set_error("cuLaunchKernel(" + e.entry + ", grid=" + std::to_string(now) + ", block=" + std::to_string(e.group[0]) + "x" +
			          std::to_string(e.group[1]) + "x" + std::to_string(e.group[2]) + "): " + drv_err(r));
			return 1;
		}
		g_state.launches++;
		done += now;
	}
	return 0;
}

int tfcuda_dispatch(const TFDispatchInfo* info) {
	uint64_t ptrs[256];
	size_t n = info->read_write_count + info->read_only_count;
	if (n > 256) { set_error("tfcuda_dispatch: too many buffers"); return 1; }
	for (size_t i = 0; i < info->read_write_count; i++) ptrs[i] = dptr_of(info->read_write_tensors[i].buffer);
	for (size_t i = 0; i < info->read_only_count; i++) ptrs[info->read_write_count + i] = dptr_of(info->read_only_tensors[i].buffer);
	if (g_profile_on) {
		double bytes = 0;
		for (size_t i = 0; i < info->read_write_count; i++) {
			size_t words = 1;
			for (size_t d = 0; d < info->read_write_tensors[i].dim; d++) words *= info->read_write_tensors[i].shape[d];
			bytes += 4.0 * words;
		}
		tfcuda_profile_add_bytes(info->kernel_id, bytes);
	}
	return tfcuda_launch(info->kernel_id, ptrs, n, info->variables, info->variable_count, info->work_group_count);
}

uint64_t tfcuda_launch_count(void) { return g_state.launches; }

int tfcuda_profile_enable(int on) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (!on && g_profile_on) profile_resolve();
	g_profile_on = on != 0;
	return 0;
}

int tfcuda_profile_reset(void) {
	if (g_state.initialized) profile_resolve();
	g_profile.clear();
	return 0;
}

int tfcuda_profile_add_bytes(size_t kernel_id, double bytes) {
	if (!g_profile_on) return 0;
	if (kernel_id >= g_kernels.size()) { set_error("tfcuda_profile_add_bytes: unknown kernel"); return 1; }
	g_profile[g_kernels[kernel_id].entry].bytes += bytes;
	return 0;
}

size_t tfcuda_profile_records(TFCudaProfileRecord* out, size_t capacity) {
	if (g_state.initialized) profile_resolve();
	size_t i = 0;
	for (auto& kv : g_profile) {
		if (out && i < capacity) {
			memset(&out[i], 0, sizeof(out[i]));
			strncpy(out[i].name, kv.first.c_str(), sizeof(out[i].name) - 1);
			out[i].launches = kv.second.launches;
			out[i].total_ms = kv.second.total_ms;
			out[i].bytes = kv.second.bytes;
		}
		i++;
	}
	return i;
}

void* tfcuda_host_alloc(size_t bytes) {
	void* p = nullptr;
	cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
	if (e != cudaSuccess) {
		set_error("tfcuda_host_alloc: " + cuda_err(e));
		(void)cudaGetLastError();
		return nullptr;
	}
	return p;
}

int tfcuda_host_free(void* p) {
	if (p) TFCUDA_CHECK(cudaFreeHost(p));
	return 0;
}

int tfcuda_timer_begin(void) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	TFCUDA_CHECK(cudaEventRecord(g_state.ev_begin, g_state.stream));
	return 0;
}

int tfcuda_timer_end(float* ms) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	TFCUDA_CHECK(cudaEventRecord(g_state.ev_end, g_state.stream));
	TFCUDA_CHECK(cudaEventSynchronize(g_state.ev_end));
	TFCUDA_CHECK(cudaEventElapsedTime(ms, g_state.ev_begin, g_state.ev_end));
	return 0;
}

}  // extern "C"

// tiny utility kernel: 32-bit fill (tf.write and buffer clears)
__global__ void tfcuda_fill32_kernel(uint32_t* p, uint32_t v, size_t n) {
	size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
	size_t stride = (size_t)gridDim.x * blockDim.x;
	for (; i < n; i += stride) p[i] = v;
}

extern "C" int tfcuda_memset32(uint64_t dst, uint32_t value, size_t words) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (words == 0) return 0;
	size_t blocks = std::min<size_t>((words + 255) / 256, (size_t)g_state.sm_count * 8);
	tfcuda_fill32_kernel<<<(unsigned)blocks, 256, 0, g_state.stream>>>(reinterpret_cast<uint32_t*>(dst), value, words);
	return check_launch("tfcuda_memset32");
}
