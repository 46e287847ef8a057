"""ctypes binding of the C-ABI in include/tfcuda.h (libtfcuda.so).  One declaration per exported symbol;
`EXPORTS` is what tests/test_abi.py checks against the header."""
import ctypes as C
import os

from . import LIB_PATH, MODULE_DIR

u64, sz, i32, u32, f32 = C.c_uint64, C.c_size_t, C.c_int, C.c_uint32, C.c_float
TF_FLOAT, TF_UINT, TF_INT, TF_BOOL, TF_NONE = 0, 1, 2, 3, 4
RED = {"sum": 0, "max": 1, "min": 2, "mean": 3, "norm": 4, "prod": 5, "any": 6, "all": 7}


class TFDataFormat(C.Structure):
    _fields_ = [("type", C.c_int), ("size", sz)]


class TFBuffer(C.Structure):
    _fields_ = [("size", sz), ("used_size", sz), ("time_since_used", sz), ("up_to_date", C.c_bool), ("read_only", C.c_bool),
                ("name", C.c_char_p)]


class TFTensor(C.Structure):
    _fields_ = [("buffer", C.POINTER(TFBuffer)), ("format", TFDataFormat), ("dim", sz), ("shape", C.POINTER(sz))]


class TFDispatchInfo(C.Structure):
    _fields_ = [("kernel_id", sz), ("read_write_count", sz), ("read_write_tensors", C.POINTER(TFTensor)), ("read_only_count", sz),
                ("read_only_tensors", C.POINTER(TFTensor)), ("variable_count", sz), ("variables", C.POINTER(u32)), ("work_group_count", sz)]


class TFRuntime(C.Structure):
    _fields_ = [("alloc", C.c_void_p), ("dealloc", C.c_void_p), ("readback", C.c_void_p), ("writeback", C.c_void_p),
                ("dispatch", C.c_void_p), ("region", C.c_void_p), ("custom_data", C.c_void_p)]


class TFCudaKernelSource(C.Structure):
    _fields_ = [("kernel_id", sz), ("entry", C.c_char_p), ("source", C.c_char_p), ("group", C.c_uint * 3), ("n_mem", C.c_uint),
                ("n_var", C.c_uint), ("library_op", C.c_uint)]


class TFCudaGraphStats(C.Structure):
    _fields_ = [("enabled", C.c_int), ("replays", u64), ("exact_hits", u64), ("patched", u64), ("instantiated", u64), ("eager_launches", u64), ("host_us", C.c_double)]


class TFCudaProfileRecord(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("launches", u64), ("total_ms", C.c_double), ("bytes", C.c_double)]


# name -> (restype, argtypes)
EXPORTS = {
    "tfcuda_init": (i32, [i32]),
    "tfcuda_is_initialized": (i32, []),
    "tfcuda_shutdown": (i32, []),
    "tfcuda_last_error": (C.c_char_p, []),
    "tfcuda_device_sm_count": (i32, []),
    "tfcuda_device_index": (i32, []),
    "tfcuda_device_name": (C.c_char_p, []),
    "tfcuda_stream": (C.c_void_p, []),
    "tfcuda_sync": (i32, []),
    "tfcuda_runtime": (TFRuntime, []),
    "tfcuda_buffer_create": (C.POINTER(TFBuffer), [sz]),
    "tfcuda_buffer_destroy": (None, [C.POINTER(TFBuffer)]),
    "tfcuda_buffer_device_ptr": (u64, [C.POINTER(TFBuffer)]),
    "tfcuda_buffer_write": (i32, [C.POINTER(TFBuffer), sz, C.c_void_p, sz]),
    "tfcuda_buffer_read": (i32, [C.POINTER(TFBuffer), sz, C.c_void_p, sz]),
    "tfcuda_memcpy_h2d": (i32, [u64, C.c_void_p, sz]),
    "tfcuda_memcpy_d2h": (i32, [C.c_void_p, u64, sz]),
    "tfcuda_memcpy_d2d": (i32, [u64, u64, sz]),
    "tfcuda_memcpy_h2d_async": (i32, [u64, C.c_void_p, sz]),
    "tfcuda_wait_uploads": (i32, []),
    "tfcuda_memcpy_d2h_async": (i32, [C.c_void_p, u64, sz]),
    "tfcuda_copy_sync": (i32, []),
    "tfcuda_downloads_issued": (u64, []),
    "tfcuda_downloads_done": (u64, []),
    "tfcuda_memset32": (i32, [u64, u32, sz]),
    "tfcuda_malloc": (u64, [sz]),
    "tfcuda_free": (i32, [u64]),
    "tfcuda_pool_allocated_words": (sz, []),
    "tfcuda_pool_unused_words": (sz, []),
    "tfcuda_pool_driver_calls": (u64, []),
    "tfcuda_prelude": (C.c_char_p, []),
    "tfcuda_nvrtc_check": (i32, [C.c_char_p, C.c_char_p]),
    "tfcuda_compile_kernels": (i32, [C.POINTER(TFCudaKernelSource), sz, C.c_char_p]),
    "tfcuda_cache_dir": (C.c_char_p, []),
    "tfcuda_launch": (i32, [sz, C.POINTER(u64), sz, C.POINTER(u32), sz, sz]),
    "tfcuda_dispatch": (i32, [C.POINTER(TFDispatchInfo)]),
    "tfcuda_graph_begin": (i32, []),
    "tfcuda_graph_end": (i32, []),
    "tfcuda_graph_stats": (i32, [C.POINTER(TFCudaGraphStats)]),
    "tfcuda_launch_count": (u64, []),
    "tfcuda_timer_begin": (i32, []),
    "tfcuda_timer_end": (i32, [C.POINTER(f32)]),
    "tfcuda_profile_enable": (i32, [i32]),
    "tfcuda_profile_reset": (i32, []),
    "tfcuda_profile_add_bytes": (i32, [sz, C.c_double]),
    "tfcuda_profile_records": (sz, [C.POINTER(TFCudaProfileRecord), sz]),
    "tfcuda_host_alloc": (C.c_void_p, [sz]),
    "tfcuda_host_free": (i32, [C.c_void_p]),
    "tfcuda_reduce": (i32, [u64, u64, sz, sz, sz, i32, i32]),
    "tfcuda_prefix_sum": (i32, [u64, u64, sz, sz, sz, i32]),
    "tfcuda_radix_sort_temp_words": (sz, [sz]),
    "tfcuda_radix_sort": (i32, [u64, u64, u64, u64, sz, i32, i32, u64]),
    "tfcuda_scatter_add": (i32, [u64, u64, u64, sz, sz, i32]),
    "tfcuda_matmul": (i32, [u64, u64, u64, sz, sz, sz, sz, i32]),
    "tfcuda_matmul_tn": (i32, [u64, u64, u64, sz, sz, sz]),
    "tfcuda_matmul_rows_supported": (i32, [sz, sz, sz]),
    "tfcuda_matmul_rows": (i32, [u64, u64, u64, sz, sz, sz]),
    "tfcuda_nbody_step": (i32, [u64, u64, u64, u64, sz, f32, f32]),
    "tfcuda_comm_unique_id": (i32, [C.c_void_p]),
    "tfcuda_comm_init": (i32, [C.c_void_p, i32, i32]),
    "tfcuda_comm_allreduce_sum_f32": (i32, [u64, sz, f32]),
    "tfcuda_comm_destroy": (i32, []),
    "tfcuda_peer_export": (i32, [C.c_void_p]),
    "tfcuda_peer_init": (i32, [C.c_void_p, i32, i32]),
    "tfcuda_peer_ready": (i32, []),
    "tfcuda_peer_max_count": (sz, []),
    "tfcuda_peer_allreduce_sum_f32": (i32, [u64, sz, f32]),
    "tfcuda_peer_destroy": (i32, []),
}

_lib = None


def lib():
    """The loaded library with typed entry points.  Raises OSError when it was never built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError(f"{LIB_PATH} not built; run tensorfrost_b200/build_lib.sh — there is no fallback implementation")
        # ONE runtime per process: the CUDA-enabled TensorFrost module resolves libtfcuda.so next to itself (RPATH $ORIGIN), so when
        # that copy exists it is the file to open here too - the dynamic loader then maps a single library (one stream, one pool),
        # whichever of the two is loaded first
        beside_module = os.path.join(MODULE_DIR, "TensorFrost", "libtfcuda.so")
        path = beside_module if os.path.exists(beside_module) else LIB_PATH
        l = C.CDLL(path, mode=C.RTLD_GLOBAL)
        l.path = path
        for name, (res, args) in EXPORTS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class TfcudaError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        raise TfcudaError(f"{what}: {lib().tfcuda_last_error().decode(errors='replace')}")


def init(device=-1):
    check(lib().tfcuda_init(device), "tfcuda_init")


class DeviceArray:
    """A device allocation holding a numpy array's words (test / bench convenience over tfcuda_malloc)."""

    def __init__(self, array=None, words=None):
        import numpy as np
        self.np = np
        if array is not None:
            array = np.ascontiguousarray(array)
            assert array.dtype.itemsize == 4
            self.shape, self.dtype, self.words = array.shape, array.dtype, array.size
        else:
            self.shape, self.dtype, self.words = (words,), np.dtype(np.uint32), words
        self.ptr = lib().tfcuda_malloc(max(self.words, 1) * 4)
        if not self.ptr:
            raise TfcudaError(lib().tfcuda_last_error().decode())
        if array is not None and array.size:
            check(lib().tfcuda_memcpy_h2d(self.ptr, array.ctypes.data, array.size * 4), "h2d")

    def get(self, dtype=None, shape=None):
        out = self.np.empty(self.shape if shape is None else shape, dtype=self.dtype if dtype is None else dtype)
        if out.size:
            check(lib().tfcuda_memcpy_d2h(out.ctypes.data, self.ptr, out.size * 4), "d2h")
        return out

    def free(self):
        if self.ptr:
            lib().tfcuda_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
