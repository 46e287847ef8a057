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
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "tfcuda_internal.h"

namespace tfcuda {

static State g_state;
static thread_local std::string g_error;
static std::string g_error_shared;
static std::mutex g_error_mutex;

static void flush_recorded();

// Library kernels (reduce.cu, matmul*.cu, ...) begin with state(): launches recorded for graph replay (below) are issued first, so
// everything that touches the stream directly stays in program order behind them.
State& state() {
	flush_recorded();
	return g_state;
}

// the runtime stream for work issued directly (copies, memsets, allocations, events): recorded launches go first
static cudaStream_t S() {
	flush_recorded();
	return g_state.stream;
}

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
// zero-filled bands (zeroed once, when the buffer is created; emitted kernels never write out of range): reads that overshoot the
// BUFFER by less than the band see zeros - the value a fresh heap gives the reference - independent of what ran before.  (The
// reference pool above may hand a tensor a recycled buffer up to 16x larger than it needs, TensorMemory.cpp:186-213: an overshoot past
// the tensor but inside such a buffer reads the previous tenant's data, on this backend as on the reference's.)
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
	if (g_scratch) cudaFreeAsync(g_scratch, S());
	g_scratch = nullptr;
	g_scratch_bytes = 0;
	size_t want = (bytes + (bytes >> 3) + 0xfffff) & ~size_t(0xfffff);
	void* p = nullptr;
	cudaError_t e = cudaMallocAsync(&p, want, S());
	if (e != cudaSuccess) {
		// parked tensor blocks / the driver pool may hold what the scratch needs: release them and ask for the exact size
		(void)cudaGetLastError();
		flush_block_cache();
		cudaStreamSynchronize(S());
		cudaMemPool_t mp;
		if (cudaDeviceGetDefaultMemPool(&mp, g_state.device) == cudaSuccess) cudaMemPoolTrimTo(mp, 0);
		(void)cudaGetLastError();
		want = bytes;
		e = cudaMallocAsync(&p, want, S());
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
			cudaFreeAsync(p, S());
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
			cudaMemsetAsync(static_cast<char*>(p) + guard + words * sizeof(uint32_t), 0, payload - words * sizeof(uint32_t), S());
	} else {
		cudaError_t e = cudaMallocAsync(&p, total, S());
		g_blocks.driver_calls++;
		if (e != cudaSuccess) {
			// parked blocks / the driver pool may be holding memory another size could use: release and retry once
			flush_block_cache();
			cudaStreamSynchronize(S());
			cudaMemPool_t mp;
			if (cudaDeviceGetDefaultMemPool(&mp, g_state.device) == cudaSuccess) cudaMemPoolTrimTo(mp, 0);
			(void)cudaGetLastError();
			e = cudaMallocAsync(&p, total, S());
			g_blocks.driver_calls++;
		}
		if (e != cudaSuccess) {
			std::string m = "tfcuda: device allocation of " + std::to_string(words * 4) + " bytes failed: " + cuda_err(e);
			set_error(m);
			throw std::runtime_error(m);
		}
		if (guard) {
			cudaMemsetAsync(p, 0, guard, S());
			cudaMemsetAsync(static_cast<char*>(p) + guard + words * sizeof(uint32_t), 0, total - guard - words * sizeof(uint32_t), S());
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
			cudaFreeAsync(p, S());
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
	                             cudaMemcpyDeviceToHost, S()));
	TFCUDA_THROW(cudaStreamSynchronize(S()));
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

// On-disk compile cache (cubins here, host-program libraries in the backend glue).  Default location: $TFCUDA_CACHE_DIR, else
// $XDG_CACHE_HOME/tfcuda, else ~/.cache/tfcuda - never a predictable path under /tmp.  The directory must belong to this user and be
// closed to everyone else (another user could otherwise plant device code that cuModuleLoadData would run in this process);
// when that cannot be established the cache is disabled.
static std::string make_cache_dir() {
	if (getenv("TFCUDA_NO_CACHE") != nullptr) return "";
	std::string dir;
	if (const char* env = getenv("TFCUDA_CACHE_DIR")) {
		dir = env;
	} else if (const char* xdg = getenv("XDG_CACHE_HOME")) {
		if (*xdg) dir = std::string(xdg) + "/tfcuda";
	}
	if (dir.empty()) {
		const char* home = getenv("HOME");
		if (!home || !*home) return "";
		std::string base = std::string(home) + "/.cache";
		mkdir(base.c_str(), 0700);
		dir = base + "/tfcuda";
	}
	mkdir(dir.c_str(), 0700);
	struct stat st;
	if (lstat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode) || st.st_uid != getuid() || (st.st_mode & 077) != 0) return "";
	return dir;
}

static const std::string& cache_dir() {
	static const std::string dir = make_cache_dir();
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

// Cached cubin file = { magic, key length, second hash of the key, payload }: a 64-bit FNV name alone would accept a colliding or
// truncated file.
struct CubinHeader {
	uint64_t magic, key_size, key_hash2, payload_size;
};
static const uint64_t kCubinMagic = 0x3130554355434654ull;  // "TFCUCU01"

static void compile_chunk(Chunk& c, const std::vector<std::string>& opts, bool use_cache) {
	std::string key_src = c.source;
	for (auto& o : opts) key_src += "\x01" + o;
	int maj = 0, min = 0;
	nvrtcVersion(&maj, &min);
	key_src += "\x01nvrtc" + std::to_string(maj) + "." + std::to_string(min);
	char name[64];
	snprintf(name, sizeof(name), "%016llx.cubin", (unsigned long long)fnv1a(key_src));
	use_cache = use_cache && !cache_dir().empty();
	const uint64_t hash2 = fnv1a(key_src, 0x9e3779b97f4a7c15ull);
	std::string path = use_cache ? cache_dir() + "/" + name : std::string();
	std::string file;
	if (use_cache && read_file(path, file) && file.size() > sizeof(CubinHeader)) {
		CubinHeader h;
		memcpy(&h, file.data(), sizeof(h));
		if (h.magic == kCubinMagic && h.key_size == key_src.size() && h.key_hash2 == hash2 && h.payload_size == file.size() - sizeof(h)) {
			c.cubin = file.substr(sizeof(h));
			c.ok = true;
			c.from_cache = true;
			return;
		}
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
	if (c.ok && use_cache) {
		CubinHeader h{kCubinMagic, key_src.size(), hash2, c.cubin.size()};
		write_file_atomic(path, std::string(reinterpret_cast<const char*>(&h), sizeof(h)) + c.cubin);
	}
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
	cudaEventRecord(smp.a, S());
	pe.pending.push_back(smp);
}

void profile_end(const char* name, double bytes) {
	if (!g_profile_on) return;
	ProfileEntry& pe = g_profile[name];
	if (pe.pending.empty()) return;
	cudaEventRecord(pe.pending.back().b, S());
	pe.launches++;
	pe.bytes += bytes;
}

static void profile_resolve() {
	cudaStreamSynchronize(S());
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


// ------------------------------------------------------------------------------------------------
// Launch recorder + graph replay ("whole-program CUDA graph").
//
// A compiled TensorFrost program is a HOST function that allocates, dispatches and frees in a fixed order
// (Backend/Backend.cpp:137-185 calls it once per prog(...)); on the reference backends every tf.dispatch is one driver launch.
// Many programs are chains of short kernels over L2-resident data (the 32 multigrid sweeps of the fluid step take 3.6 us each,
// almost all of it launch latency).  Between tfcuda_graph_begin and tfcuda_graph_end (the backend glue brackets every program
// execution) tfcuda_launch therefore only RECORDS {function, grid, block, argument block}.  The list is issued when the program
// ends - or earlier, the moment anything else needs the stream (a tf.read, a copy, a library kernel, a driver allocation: every
// such path goes through S() / state()) - as ONE cudaGraphLaunch of a linear kernel-node chain:
//   * exact hit    (same kernels, same argument bytes as an executable graph built earlier: the pool hands out the same addresses
//                   every step, or alternates between two sets when outputs are fed back)  -> cuGraphLaunch, nothing else;
//   * second sight (same kernels and the same argument bytes as a recent execution that was launched eagerly: the chain repeats)
//                                                                                         -> build + instantiate, then replay
//                   (at most kExecsPerShape executable graphs per kernel sequence, least recently used evicted);
//   * first sight  (arguments never seen: a training step whose buffers come back in a different order every time)
//                                                                                         -> launch eagerly, remember the hash.
// A chain therefore pays for a graph only once it has proven to repeat: measured at 8 GPUs on the NCA step (whose chains never
// repeat), instantiating on every miss cost 10 ms of a 73 ms step under host contention; with this rule the step takes 63 ms.
// The argument bytes are baked into the nodes, so a replay launches exactly what eager execution would have launched, in the
// same order; results are bit-identical (tests/test_graph_replay_gpu.py).  Lists shorter than kMinGraphOps, and lists without at least
// kMinGraphOps launch-latency-sized kernels, are launched eagerly.
// ------------------------------------------------------------------------------------------------
struct RecOp {
	CUfunction fn;
	unsigned grid;
	unsigned block[3];
	uint32_t arg_offset, arg_bytes;
};
struct GraphExec {
	CUgraph graph = nullptr;
	CUgraphExec exec = nullptr;
	std::vector<CUgraphNode> nodes;
	std::vector<unsigned char> args;
	uint64_t args_hash = 0;
	uint64_t last_used = 0;
};
struct GraphShape {
	std::vector<RecOp> ops;  // arg_offset / arg_bytes included: equal shapes have equal layouts
	std::vector<GraphExec> execs;
	// argument hashes of the last executions that were launched eagerly.  As many as executable graphs are kept (kExecsPerShape): a chain
	// whose arguments cycle with a longer period would be instantiated, evicted before its next turn and instantiated again, for ever.
	uint64_t seen[4] = {0, 0, 0, 0};
	unsigned seen_at = 0;
};
struct Recorder {
	int depth = 0;            // tfcuda_graph_begin nesting
	bool enabled = false;     // TFCUDA_GRAPH (default on) and the driver entry points resolved
	bool pdl = false;         // TFCUDA_PDL=1: kernel nodes are chained with PROGRAMMATIC edges (kernels start with griddepcontrol.wait)
	bool flushing = false;
	std::vector<RecOp> ops;
	std::vector<unsigned char> args;
	std::unordered_map<uint64_t, GraphShape> shapes;
	uint64_t tick = 0;
	uint64_t replays = 0, exact_hits = 0, patched = 0, instantiated = 0, eager = 0;
	double host_us = 0.0;     // host time spent issuing recorded chains (hashing, graph upkeep, launches)
	std::string error;        // first failure of a deferred launch; reported by the next tfcuda_launch / tfcuda_graph_end / tfcuda_sync
};
static Recorder g_rec;
static const size_t kMinGraphOps = 4;
// below ~1 M threads a kernel on 148 SMs is over in a few microseconds.  (A kernel the emitter coarsened carries 4 elements per
// thread: one of up to 2^22 elements counts as short here, which it is - the 2048^2 fluid stencils take 11-13 us with lanes.)
static const uint64_t kSmallKernelThreads = 1ull << 20;
static const size_t kExecsPerShape = 4;
static const size_t kMaxShapes = 256;

static uint64_t hash_bytes(const void* p, size_t n, uint64_t h = 1469598103934665603ull) {
	const unsigned char* b = static_cast<const unsigned char*>(p);
	size_t i = 0;
	for (; i + 8 <= n; i += 8) {
		uint64_t w;
		memcpy(&w, b + i, 8);
		h = (h ^ w) * 1099511628211ull;
		h ^= h >> 29;
	}
	for (; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
	return h;
}

static bool rec_fail(const std::string& what, CUresult r) {
	if (g_rec.error.empty()) g_rec.error = what + ": " + drv_err(r);
	return false;
}

static bool launch_eager(const RecOp& op, const unsigned char* args) {
	void* params[1] = {const_cast<unsigned char*>(args) + op.arg_offset};
	CUresult r = g_state.drv.LaunchKernel(op.fn, op.grid, 1, 1, op.block[0], op.block[1], op.block[2], 0, (CUstream)g_state.stream, params, nullptr);
	if (r != CUDA_SUCCESS) return rec_fail("cuLaunchKernel (deferred)", r);
	return true;
}

static void fill_node_params(CUDA_KERNEL_NODE_PARAMS& kp, const RecOp& op, void** params) {
	memset(&kp, 0, sizeof(kp));
	kp.func = op.fn;
	kp.gridDimX = op.grid; kp.gridDimY = 1; kp.gridDimZ = 1;
	kp.blockDimX = op.block[0]; kp.blockDimY = op.block[1]; kp.blockDimZ = op.block[2];
	kp.sharedMemBytes = 0;
	kp.kernelParams = params;
	kp.extra = nullptr;
}

static void destroy_exec(GraphExec& ge) {
	if (ge.exec) g_state.drv.GraphExecDestroy(ge.exec);
	if (ge.graph) g_state.drv.GraphDestroy(ge.graph);
	ge = GraphExec();
}

static bool build_exec(GraphExec& ge, const std::vector<RecOp>& ops, const std::vector<unsigned char>& args) {
	DriverApi& d = g_state.drv;
	CUresult r = d.GraphCreate(&ge.graph, 0);
	if (r != CUDA_SUCCESS) return rec_fail("cuGraphCreate", r);
	ge.nodes.resize(ops.size());
	const bool pdl = g_rec.pdl && d.GraphAddDependencies_v2 != nullptr;
	for (size_t i = 0; i < ops.size(); i++) {
		void* params[1] = {const_cast<unsigned char*>(args.data()) + ops[i].arg_offset};
		CUDA_KERNEL_NODE_PARAMS kp;
		fill_node_params(kp, ops[i], params);
		r = d.GraphAddKernelNode(&ge.nodes[i], ge.graph, (i && !pdl) ? &ge.nodes[i - 1] : nullptr, (i && !pdl) ? 1 : 0, &kp);
		if (r != CUDA_SUCCESS) {
			destroy_exec(ge);
			return rec_fail("cuGraphAddKernelNode", r);
		}
	}
	if (pdl && ops.size() > 1) {
		// programmatic edges: node i+1 may be launched once every CTA of node i has executed griddepcontrol.launch_dependents (the
		// first statement of every emitted kernel under TFCUDA_PDL=1); its own griddepcontrol.wait holds it until node i has completed
		// and flushed, so the order of memory effects is the chain's.  What overlaps is the launch latency and the drain of node i.
		const size_t m = ops.size() - 1;
		std::vector<CUgraphEdgeData> edges(m);
		memset(edges.data(), 0, m * sizeof(CUgraphEdgeData));
		for (size_t i = 0; i < m; i++) {
			edges[i].from_port = CU_GRAPH_KERNEL_NODE_PORT_PROGRAMMATIC;
			edges[i].to_port = 0;
			edges[i].type = CU_GRAPH_DEPENDENCY_TYPE_PROGRAMMATIC;
		}
		r = d.GraphAddDependencies_v2(ge.graph, ge.nodes.data(), ge.nodes.data() + 1, edges.data(), m);
		if (r != CUDA_SUCCESS) {
			destroy_exec(ge);
			return rec_fail("cuGraphAddDependencies (programmatic)", r);
		}
	}
	r = d.GraphInstantiate(&ge.exec, ge.graph, 0);
	if (r != CUDA_SUCCESS) {
		destroy_exec(ge);
		return rec_fail("cuGraphInstantiate", r);
	}
	ge.args = args;
	return true;
}

static void flush_recorded() {
	Recorder& R = g_rec;
	if (R.ops.empty() || R.flushing) return;
	R.flushing = true;
	std::vector<RecOp> ops;
	std::vector<unsigned char> args;
	ops.swap(R.ops);
	args.swap(R.args);
	const size_t n = ops.size();
	bool done = false;
	const auto t_begin = std::chrono::steady_clock::now();
	// What a graph saves is launch latency, so it only pays for chains of SHORT kernels (the 3-6 us multigrid sweeps of the fluid
	// step); a chain of kernels with millions of threads each (an NCA training step) is device-bound whatever the launch path is,
	// and graph upkeep would only add host work.  A chain qualifies when at least kMinGraphOps of its kernels are small.
	size_t small_ops = 0;
	for (const RecOp& op : ops) small_ops += ((uint64_t)op.grid * op.block[0] * op.block[1] * op.block[2] < kSmallKernelThreads) ? 1 : 0;
	if (n >= kMinGraphOps && small_ops >= kMinGraphOps && R.error.empty()) {
		const uint64_t shape_hash = hash_bytes(ops.data(), n * sizeof(RecOp));
		const uint64_t args_hash = hash_bytes(args.data(), args.size());
		if (R.shapes.size() >= kMaxShapes && !R.shapes.count(shape_hash)) {
			for (auto& kv : R.shapes)
				for (GraphExec& ge : kv.second.execs) destroy_exec(ge);
			R.shapes.clear();
		}
		GraphShape& gs = R.shapes[shape_hash];
		bool same_shape = gs.ops.size() == n && memcmp(gs.ops.data(), ops.data(), n * sizeof(RecOp)) == 0;
		if (!same_shape) {  // first time, or a 64-bit collision: start over for this key
			for (GraphExec& ge : gs.execs) destroy_exec(ge);
			gs.execs.clear();
			gs.ops = ops;
		}
		GraphExec* use = nullptr;
		for (GraphExec& ge : gs.execs)
			if (ge.args_hash == args_hash && ge.args.size() == args.size() && memcmp(ge.args.data(), args.data(), args.size()) == 0) {
				use = &ge;
				R.exact_hits++;
				break;
			}
		if (!use) {
			bool seen_before = false;
			for (uint64_t h : gs.seen) seen_before |= (h == args_hash);
			if (seen_before) {
				// these exact arguments were launched (eagerly) not long ago: the chain repeats, a graph pays off from now on
				if (gs.execs.size() >= kExecsPerShape) {
					size_t lru = 0;
					for (size_t i = 1; i < gs.execs.size(); i++)
						if (gs.execs[i].last_used < gs.execs[lru].last_used) lru = i;
					destroy_exec(gs.execs[lru]);
					gs.execs.erase(gs.execs.begin() + (long)lru);
				}
				GraphExec ge;
				if (build_exec(ge, ops, args)) {
					ge.args_hash = args_hash;
					gs.execs.push_back(std::move(ge));
					use = &gs.execs.back();
					R.instantiated++;
				}
			} else {
				gs.seen[gs.seen_at++ % 4] = args_hash;
			}
		}
		if (use) {
			CUresult r = g_state.drv.GraphLaunch(use->exec, (CUstream)g_state.stream);
			if (r == CUDA_SUCCESS) {
				use->last_used = ++R.tick;
				R.replays++;
				done = true;
			} else {
				rec_fail("cuGraphLaunch", r);
			}
		}
	}
	if (!done) {
		for (size_t i = 0; i < n; i++)
			if (!launch_eager(ops[i], args.data())) break;
		R.eager += n;
	}
	g_state.launches += n;
	R.host_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count();
	R.flushing = false;
}

static void recorder_reset() {
	for (auto& kv : g_rec.shapes)
		for (GraphExec& ge : kv.second.execs) destroy_exec(ge);
	g_rec = Recorder();
}

}  // namespace tfcuda

using namespace tfcuda;

// copy engines (tfcuda_memcpy_*_async below) and the staging ring of small uploads (tfcuda_memcpy_h2d)
static unsigned char* g_ring = nullptr;
static size_t g_ring_at = 0;
static const size_t kRingBytes = 4u << 20, kRingMaxCopy = 64u << 10;
static cudaStream_t g_up_stream = nullptr, g_down_stream = nullptr;
static cudaEvent_t g_up_event = nullptr, g_down_event = nullptr, g_order_event = nullptr;
static bool g_up_pending = false;
// Completion of downloads is tracked with one event per download, polled by tfcuda_downloads_done().  (A cudaLaunchHostFunc callback
// per download was measured on the B200: it stalls the copy stream for ~0.28 ms each - downloads of 4 x 16 MB took 2.32 ms instead of
// 1.21 ms and the end-to-end fluid step 2.56 ms instead of 1.55 ms.)
static uint64_t g_down_issued = 0, g_down_done = 0;
static std::vector<std::pair<uint64_t, cudaEvent_t>> g_down_pending;  // (ticket, event), oldest first
static std::vector<cudaEvent_t> g_down_event_pool;

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

const char* tfcuda_cache_dir(void) { return cache_dir().c_str(); }

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
	{
		// graph replay of recorded launches: on unless TFCUDA_GRAPH=0; needs the driver's graph entry points
		const char* g = getenv("TFCUDA_GRAPH");
		bool want = !(g && atoi(g) == 0);
		auto opt = [](const char* sym, auto& fn) {
			void* p = nullptr;
			cudaDriverEntryPointQueryResult q;
			if (cudaGetDriverEntryPoint(sym, &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && p) {
				fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
				return true;
			}
			(void)cudaGetLastError();
			return false;
		};
		bool have = opt("cuGraphCreate", d.GraphCreate) & opt("cuGraphAddKernelNode", d.GraphAddKernelNode) &
		            opt("cuGraphInstantiateWithFlags", d.GraphInstantiate) & opt("cuGraphLaunch", d.GraphLaunch) &
		            opt("cuGraphExecKernelNodeSetParams", d.GraphExecKernelNodeSetParams) & opt("cuGraphExecDestroy", d.GraphExecDestroy) &
		            opt("cuGraphDestroy", d.GraphDestroy);
		g_rec.enabled = want && have;
		const char* pdl_env = getenv("TFCUDA_PDL");
		g_rec.pdl = pdl_env != nullptr && atoi(pdl_env) != 0 && opt("cuGraphAddDependencies", d.GraphAddDependencies_v2);
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
	cudaStreamSynchronize(S());
	recorder_reset();
	for (auto& kv : g_pool.free_lists)
		for (Buffer* b : kv.second) destroy_buffer(b);
	g_pool.free_lists.clear();
	g_pool.unused_words = 0;
	flush_block_cache();
	if (g_scratch) cudaFreeAsync(g_scratch, S());
	g_scratch = nullptr;
	g_scratch_bytes = 0;
	for (CUmodule m : g_modules) g_state.drv.ModuleUnload(m);
	g_modules.clear();
	g_kernels.clear();
	cudaEventDestroy(g_state.ev_begin);
	cudaEventDestroy(g_state.ev_end);
	cudaFreeHost(g_state.pinned_word);
	if (g_ring) cudaFreeHost(g_ring);
	g_ring = nullptr;
	g_ring_at = 0;
	if (g_up_stream) {
		cudaStreamDestroy(g_up_stream);
		cudaStreamDestroy(g_down_stream);
		cudaEventDestroy(g_up_event);
		cudaEventDestroy(g_down_event);
		cudaEventDestroy(g_order_event);
		for (auto& pe : g_down_pending) cudaEventDestroy(pe.second);
		for (cudaEvent_t e : g_down_event_pool) cudaEventDestroy(e);
		g_down_pending.clear();
		g_down_event_pool.clear();
		g_up_stream = g_down_stream = nullptr;
		g_up_pending = false;
	}
	cudaStreamDestroy(g_state.stream);
	g_state = State();
	return 0;
}

int tfcuda_device_sm_count(void) { return g_state.sm_count; }
int tfcuda_device_index(void) { return g_state.device; }
const char* tfcuda_device_name(void) { return g_state.device_name.c_str(); }
void* tfcuda_stream(void) { return g_state.initialized ? S() : nullptr; }

int tfcuda_sync(void) {
	if (!g_state.initialized) {
		set_error("tfcuda_sync: not initialised");
		return 1;
	}
	TFCUDA_CHECK(cudaStreamSynchronize(S()));
	if (!g_rec.error.empty()) {
		set_error("deferred launch failed: " + g_rec.error);
		g_rec.error.clear();
		return 1;
	}
	return 0;
}

int tfcuda_graph_begin(void) {
	if (!g_state.initialized) { set_error("tfcuda_graph_begin: not initialised"); return 1; }
	g_rec.depth++;
	return 0;
}

int tfcuda_graph_end(void) {
	if (!g_state.initialized) { set_error("tfcuda_graph_end: not initialised"); return 1; }
	if (g_rec.depth > 0) g_rec.depth--;
	if (g_rec.depth == 0) flush_recorded();
	if (!g_rec.error.empty()) {
		set_error("deferred launch failed: " + g_rec.error);
		g_rec.error.clear();
		return 1;
	}
	return 0;
}

int tfcuda_graph_stats(TFCudaGraphStats* out) {
	if (!out) return 1;
	out->enabled = g_rec.enabled ? 1 : 0;
	out->replays = g_rec.replays;
	out->exact_hits = g_rec.exact_hits;
	out->patched = g_rec.patched;
	out->instantiated = g_rec.instantiated;
	out->eager_launches = g_rec.eager;
	out->host_us = g_rec.host_us;
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

// Small uploads (program parameters, batch indices: a few bytes to a few KB per call) are staged through a ring of page-locked
// memory and copied asynchronously.  A cudaMemcpyAsync from PAGEABLE memory first synchronises the stream, so a program fed a numpy
// scalar on every call would drain the whole pipeline once per call (one of the two stalls of a data-parallel NCA step).

int tfcuda_memcpy_h2d(uint64_t dst, const void* src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	if (bytes <= kRingMaxCopy) {
		if (!g_ring) TFCUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&g_ring), kRingBytes));
		const size_t need = (bytes + 255) & ~size_t(255);
		if (g_ring_at + need > kRingBytes) {
			TFCUDA_CHECK(cudaStreamSynchronize(S()));  // wrap-around: every copy staged so far has left the ring
			g_ring_at = 0;
		}
		memcpy(g_ring + g_ring_at, src, bytes);
		TFCUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(dst), g_ring + g_ring_at, bytes, cudaMemcpyHostToDevice, S()));
		g_ring_at += need;
		return 0;
	}
	TFCUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, S()));
	// pageable sources are consumed before the call returns; pinned ones are not, so order the host too
	TFCUDA_CHECK(cudaStreamSynchronize(S()));
	return 0;
}

int tfcuda_memcpy_d2h(void* dst, uint64_t src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	TFCUDA_CHECK(cudaMemcpyAsync(dst, reinterpret_cast<const void*>(src), bytes, cudaMemcpyDeviceToHost, S()));
	TFCUDA_CHECK(cudaStreamSynchronize(S()));
	return 0;
}

// ---- copy engines: uploads and downloads on their own streams, overlapping each other and the kernels (PCIe is full duplex) ----

static int copy_streams_init() {
	if (g_up_stream) return 0;
	TFCUDA_CHECK(cudaStreamCreateWithFlags(&g_up_stream, cudaStreamNonBlocking));
	TFCUDA_CHECK(cudaStreamCreateWithFlags(&g_down_stream, cudaStreamNonBlocking));
	TFCUDA_CHECK(cudaEventCreateWithFlags(&g_up_event, cudaEventDisableTiming));
	TFCUDA_CHECK(cudaEventCreateWithFlags(&g_down_event, cudaEventDisableTiming));
	TFCUDA_CHECK(cudaEventCreateWithFlags(&g_order_event, cudaEventDisableTiming));
	return 0;
}

int tfcuda_memcpy_h2d_async(uint64_t dst, const void* src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	if (copy_streams_init()) return 1;
	// the copy starts once everything queued on the runtime stream SO FAR has finished (the previous users of dst)
	TFCUDA_CHECK(cudaEventRecord(g_order_event, S()));
	TFCUDA_CHECK(cudaStreamWaitEvent(g_up_stream, g_order_event, 0));
	TFCUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, g_up_stream));
	TFCUDA_CHECK(cudaEventRecord(g_up_event, g_up_stream));
	g_up_pending = true;
	return 0;
}

int tfcuda_wait_uploads(void) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (!g_up_pending) return 0;
	TFCUDA_CHECK(cudaStreamWaitEvent(S(), g_up_event, 0));
	g_up_pending = false;
	return 0;
}

int tfcuda_memcpy_d2h_async(void* dst, uint64_t src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	if (copy_streams_init()) return 1;
	TFCUDA_CHECK(cudaEventRecord(g_order_event, S()));
	TFCUDA_CHECK(cudaStreamWaitEvent(g_down_stream, g_order_event, 0));
	TFCUDA_CHECK(cudaMemcpyAsync(dst, reinterpret_cast<const void*>(src), bytes, cudaMemcpyDeviceToHost, g_down_stream));
	TFCUDA_CHECK(cudaEventRecord(g_down_event, g_down_stream));
	cudaEvent_t ev;
	if (!g_down_event_pool.empty()) {
		ev = g_down_event_pool.back();
		g_down_event_pool.pop_back();
	} else {
		TFCUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
	}
	TFCUDA_CHECK(cudaEventRecord(ev, g_down_stream));
	g_down_pending.emplace_back(++g_down_issued, ev);
	return 0;
}

uint64_t tfcuda_downloads_issued(void) { return g_down_issued; }

uint64_t tfcuda_downloads_done(void) {
	size_t k = 0;
	while (k < g_down_pending.size() && cudaEventQuery(g_down_pending[k].second) == cudaSuccess) {
		g_down_done = g_down_pending[k].first;
		g_down_event_pool.push_back(g_down_pending[k].second);
		k++;
	}
	(void)cudaGetLastError();  // cudaErrorNotReady of the first unfinished event is not an error
	if (k) g_down_pending.erase(g_down_pending.begin(), g_down_pending.begin() + k);
	return g_down_done;
}

int tfcuda_copy_sync(void) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (!g_up_stream) return 0;
	TFCUDA_CHECK(cudaStreamSynchronize(g_up_stream));
	TFCUDA_CHECK(cudaStreamSynchronize(g_down_stream));
	(void)tfcuda_downloads_done();  // everything issued so far has completed: recycle the events
	return 0;
}

int tfcuda_memcpy_d2d(uint64_t dst, uint64_t src, size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	if (bytes == 0) return 0;
	TFCUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<void*>(dst), reinterpret_cast<const void*>(src), bytes, cudaMemcpyDeviceToDevice, S()));
	return 0;
}

uint64_t tfcuda_malloc(size_t bytes) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 0; }
	void* p = nullptr;
	cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 4, S());
	if (e != cudaSuccess) {
		set_error("tfcuda_malloc(" + std::to_string(bytes) + "): " + cuda_err(e));
		return 0;
	}
	return reinterpret_cast<uint64_t>(p);
}

int tfcuda_free(uint64_t ptr) {
	if (!ptr) return 0;
	TFCUDA_CHECK(cudaFreeAsync(reinterpret_cast<void*>(ptr), S()));
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
	bool use_cache = !cache_dir().empty();

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
	// block ids are `int` in the generated code (CPP.cpp:503-515) and emitted kernels state `block_id >= 0` to the compiler (prelude.cuh)
	if ((n_var ? (uint64_t)vars[n_var - 1] : 0) + (uint64_t)work_group_count > 0x7fffffffull) {
		set_error("tfcuda_launch: kernel " + std::to_string(kernel_id) + ": " + std::to_string(work_group_count) +
		          " blocks exceed the 2^31-1 block ids a generated kernel can address");
		return 1;
	}
	// argument block = { uint* mem[n_mem]; uint var[n_var]; } passed by value as the single kernel parameter
	alignas(8) unsigned char block[4096 + 8];
	size_t bytes = n_mem * 8 + n_var * 4;
	if (bytes > 4096) {
		set_error("tfcuda_launch: argument block too large");
		return 1;
	}
	memcpy(block, mem, n_mem * 8);
	memcpy(block + n_mem * 8, vars, n_var * 4);
	memset(block + bytes, 0, 8);  // struct padding: recorded argument blocks are compared byte by byte
	uint32_t* offset_word = n_var ? reinterpret_cast<uint32_t*>(block + n_mem * 8 + (n_var - 1) * 4) : nullptr;
	void* params[1] = {block};
	// grid.x is limited to 2^31-1: larger dispatches are split with the _kernel_block_offset word
	const size_t kMaxGrid = 0x7fffffffull;
	size_t done = 0;
	static const bool pdl_env = getenv("TFCUDA_PDL") != nullptr && atoi(getenv("TFCUDA_PDL")) != 0;
	const bool pdl = pdl_env && g_state.drv.LaunchKernelEx != nullptr;
	if (g_rec.depth > 0 && g_rec.enabled && !g_profile_on && (!pdl || g_rec.pdl)) {
		// inside a program execution: record, the list is replayed as one graph (see "Launch recorder")
		if (!g_rec.error.empty()) {
			set_error("deferred launch failed: " + g_rec.error);
			g_rec.error.clear();
			return 1;
		}
		while (done < work_group_count) {
			size_t now = std::min(kMaxGrid, work_group_count - done);
			if (offset_word) *offset_word = vars[n_var - 1] + (uint32_t)done;
			RecOp op;
			memset(&op, 0, sizeof(op));
			op.fn = e.fn;
			op.grid = (unsigned)now;
			for (int d = 0; d < 3; d++) op.block[d] = e.group[d];
			op.arg_offset = (uint32_t)g_rec.args.size();
			op.arg_bytes = (uint32_t)bytes;
			g_rec.args.insert(g_rec.args.end(), block, block + ((bytes + 7) & ~size_t(7)));
			g_rec.ops.push_back(op);
			done += now;
		}
		return 0;
	}
	flush_recorded();  // (nothing is pending unless the mode changed in the middle of a program)
	ProfileScope prof(e.entry.c_str());
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
	TFCUDA_CHECK(cudaEventRecord(g_state.ev_begin, S()));
	return 0;
}

int tfcuda_timer_end(float* ms) {
	if (!g_state.initialized) { set_error("tfcuda: not initialised"); return 1; }
	TFCUDA_CHECK(cudaEventRecord(g_state.ev_end, S()));
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
	tfcuda_fill32_kernel<<<(unsigned)blocks, 256, 0, S()>>>(reinterpret_cast<uint32_t*>(dst), value, words);
	return check_launch("tfcuda_memset32");
}
