"""GPU parity of the hand-written library kernels, called through the C-ABI (include/tfcuda.h) with ctypes,
against the CPU restatement of the reference algorithms (oracle/tf_oracle.py, itself pinned to reference outputs in
tests/test_oracle.py).  Bit-exact for sort / integer / scan / integer-valued atomics; fp32 tolerance stated per test."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import tf_oracle  # noqa: E402

from tensorfrost_b200 import abi  # noqa: E402

pytestmark = pytest.mark.gpu

TYPE = {np.dtype(np.float32): abi.TF_FLOAT, np.dtype(np.int32): abi.TF_INT, np.dtype(np.uint32): abi.TF_UINT}


def rel_err(got, want):
    scale = max(float(np.max(np.abs(want))), 1e-30)
    return float(np.max(np.abs(got.astype(np.float64) - want.astype(np.float64)))) / scale


# ---- reduce -------------------------------------------------------------------------------------------------
def gpu_reduce(lib, a, axis, op):
    axis = axis % a.ndim
    outer = int(np.prod(a.shape[:axis], dtype=np.int64))
    inner = int(np.prod(a.shape[axis + 1:], dtype=np.int64))
    out_shape = a.shape[:axis] + a.shape[axis + 1:]
    d_in = abi.DeviceArray(a)
    d_out = abi.DeviceArray(np.zeros(max(outer * inner, 1), a.dtype))
    abi.check(lib.tfcuda_reduce(d_in.ptr, d_out.ptr, outer, a.shape[axis], inner, abi.RED[op], TYPE[a.dtype]), "reduce")
    return d_out.get().reshape(out_shape if out_shape else (1,))


@pytest.mark.parametrize("shape,axis", [((300, 8192), -1), ((64, 1000), -1), ((1000, 37), -1), ((5, 2048, 33), 1), ((4000, 96), 0), ((3, 7), 1)])
@pytest.mark.parametrize("op", ["sum", "max", "min", "mean", "norm"])
def test_reduce_f32(tfcuda_lib, shape, axis, op):
    rng = np.random.default_rng(1)
    a = (rng.random(shape, dtype=np.float32) - 0.25).astype(np.float32)
    got = gpu_reduce(tfcuda_lib, a, axis, op)
    want = tf_oracle.reduce(a, axis, op)
    if op in ("max", "min"):
        assert np.array_equal(got, want)
    else:
        # tree vs serial fp32 summation: bound by 1e-5 of the result scale (north_star: 1e-5 relative for reductions)
        assert rel_err(got, want) <= 1e-5


@pytest.mark.parametrize("dtype", [np.int32, np.uint32])
@pytest.mark.parametrize("op", ["sum", "max", "min", "any", "all"])
def test_reduce_int_exact(tfcuda_lib, dtype, op):
    rng = np.random.default_rng(2)
    a = rng.integers(0 if dtype is np.uint32 else -1000, 1000, (257, 3001)).astype(dtype)
    if op in ("any", "all"):
        a[::3] = 0
        a[1::3] = 1
    for axis in (0, 1):
        assert np.array_equal(gpu_reduce(tfcuda_lib, a, axis, op), tf_oracle.reduce(a, axis, op))


# ---- prefix sum ----------------------------------------------------------------------------------------------
def gpu_scan(lib, a, axis):
    axis = axis % a.ndim
    outer = int(np.prod(a.shape[:axis], dtype=np.int64))
    inner = int(np.prod(a.shape[axis + 1:], dtype=np.int64))
    d_in, d_out = abi.DeviceArray(a), abi.DeviceArray(np.zeros_like(a))
    abi.check(lib.tfcuda_prefix_sum(d_in.ptr, d_out.ptr, outer, a.shape[axis], inner, TYPE[a.dtype]), "scan")
    return d_out.get()


@pytest.mark.parametrize("shape,axis", [((1,), 0), ((2047,), 0), ((2048,), 0), ((2049,), 0), ((1000003,), 0), ((7, 30000), 1), ((300, 5), 0), ((9, 100, 11), 1)])
def test_prefix_sum_exact(tfcuda_lib, shape, axis):
    rng = np.random.default_rng(3)
    a = rng.integers(-50, 50, shape, dtype=np.int32)
    assert np.array_equal(gpu_scan(tfcuda_lib, a, axis), np.cumsum(a, axis=axis, dtype=np.int32))
    u = rng.integers(0, 2 ** 32, shape, dtype=np.uint64).astype(np.uint32)  # wraps mod 2^32 like the reference
    assert np.array_equal(gpu_scan(tfcuda_lib, u, axis), np.cumsum(u, axis=axis, dtype=np.uint32))
    f = rng.integers(0, 4, shape).astype(np.float32)  # integer valued: exact in fp32 in any association
    assert np.array_equal(gpu_scan(tfcuda_lib, f, axis), tf_oracle.prefix_sum(f, axis))


def test_prefix_sum_f32_tolerance(tfcuda_lib):
    a = np.random.default_rng(4).random(500000, dtype=np.float32)
    want = tf_oracle.prefix_sum(a.astype(np.float64))  # exact-ish truth; both fp32 orders are within 1e-5 of it
    assert rel_err(gpu_scan(tfcuda_lib, a, 0), want) <= 1e-5


# ---- radix sort ----------------------------------------------------------------------------------------------
def gpu_sort(lib, keys, values=None, max_bits=32):
    n = keys.size
    d_k, d_ko = abi.DeviceArray(keys), abi.DeviceArray(np.zeros_like(keys))
    d_tmp = abi.DeviceArray(words=lib.tfcuda_radix_sort_temp_words(n))
    if values is None:
        abi.check(lib.tfcuda_radix_sort(d_k.ptr, d_ko.ptr, 0, 0, n, TYPE[keys.dtype], max_bits, d_tmp.ptr), "sort")
        assert np.array_equal(d_k.get(), keys), "input keys must be preserved"
        return d_ko.get(), None
    d_v, d_vo = abi.DeviceArray(values), abi.DeviceArray(np.zeros_like(values))
    abi.check(lib.tfcuda_radix_sort(d_k.ptr, d_ko.ptr, d_v.ptr, d_vo.ptr, n, TYPE[keys.dtype], max_bits, d_tmp.ptr), "sort")
    return d_ko.get(), d_vo.get()


def make_keys(rng, n, dtype):
    if dtype is np.float32:
        k = (rng.standard_normal(n) * 100).astype(np.float32)
        if n > 8:
            k[:4] = [0.0, -0.0, np.inf, -np.inf]
    elif dtype is np.int32:
        k = rng.integers(-2 ** 31, 2 ** 31, n, dtype=np.int64).astype(np.int32)
    else:
        k = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    if n > 16:
        k[n // 2: n // 2 + n // 4] = k[: n // 4]  # duplicates
    return k


@pytest.mark.parametrize("n", [1, 2, 31, 33, 8191, 8192, 8193, 100003, 1 << 20])
@pytest.mark.parametrize("dtype", [np.uint32, np.int32, np.float32])
def test_radix_sort_matches_stable_argsort(tfcuda_lib, n, dtype):
    rng = np.random.default_rng(n)
    keys = make_keys(rng, n, dtype)
    values = np.arange(n, dtype=np.uint32)
    order = np.argsort(tf_oracle.map_key_to_uint(keys), kind="stable")
    k, v = gpu_sort(tfcuda_lib, keys, values)
    assert np.array_equal(k.view(np.uint32), keys[order].view(np.uint32))
    assert np.array_equal(v, values[order]), "values must follow the STABLE order (LSD radix, sort.py)"
    k2, _ = gpu_sort(tfcuda_lib, keys)
    assert np.array_equal(k2.view(np.uint32), keys[order].view(np.uint32))


def test_radix_sort_equals_reference_restatement(tfcuda_lib):
    rng = np.random.default_rng(5)
    for dtype in (np.uint32, np.int32, np.float32):
        keys = make_keys(rng, 50000, dtype)
        values = rng.integers(0, 2 ** 32, 50000, dtype=np.uint64).astype(np.uint32)
        k, v = gpu_sort(tfcuda_lib, keys, values)
        rk, rv = tf_oracle.radix_sort(keys, values)  # 6-bit, 6 passes: the reference's algorithm
        assert np.array_equal(k.view(np.uint32), rk.view(np.uint32)) and np.array_equal(v, rv)


@pytest.mark.parametrize("max_bits", [1, 8, 12, 16, 20, 24])
def test_radix_sort_max_bits(tfcuda_lib, max_bits):
    rng = np.random.default_rng(max_bits)
    keys = rng.integers(0, 2 ** 32, 70001, dtype=np.uint64).astype(np.uint32)
    values = np.arange(keys.size, dtype=np.uint32)
    order = np.argsort(keys & np.uint32((1 << max_bits) - 1), kind="stable")
    k, v = gpu_sort(tfcuda_lib, keys, values, max_bits=max_bits)
    assert np.array_equal(k, keys[order]) and np.array_equal(v, values[order])


@pytest.mark.parametrize("kind", ["all_equal", "sorted", "reversed", "two_values"])
def test_radix_sort_degenerate_inputs(tfcuda_lib, kind):
    n = 300000
    keys = {"all_equal": np.full(n, 0xDEADBEEF, np.uint32), "sorted": np.arange(n, dtype=np.uint32) * 7,
            "reversed": (np.arange(n, dtype=np.uint32)[::-1] * 5).copy(),
            "two_values": (np.arange(n, dtype=np.uint32) % 2) * np.uint32(0x80000001)}[kind]
    values = np.arange(n, dtype=np.uint32)
    order = np.argsort(keys, kind="stable")
    k, v = gpu_sort(tfcuda_lib, keys, values)
    assert np.array_equal(k, keys[order]) and np.array_equal(v, values[order])


def test_radix_sort_empty(tfcuda_lib):
    assert tfcuda_lib.tfcuda_radix_sort(1, 1, 0, 0, 0, abi.TF_UINT, 32, 1) == 0


# ---- scatter add ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.int32, np.uint32, np.float32])
@pytest.mark.parametrize("bins", [1, 7, 1000, 1 << 20])
def test_scatter_add(tfcuda_lib, dtype, bins):
    rng = np.random.default_rng(bins)
    n = 400001
    idx = rng.integers(-3, bins + 3, n, dtype=np.int32)  # a few out-of-range indices: clamped like the reference
    idx[: n // 3] = bins // 2
    src = rng.integers(0, 16, n).astype(dtype)  # integer valued: any association is exact
    dst0 = rng.integers(0, 5, bins).astype(dtype)
    d_dst, d_idx, d_src = abi.DeviceArray(dst0), abi.DeviceArray(idx), abi.DeviceArray(src)
    abi.check(tfcuda_lib.tfcuda_scatter_add(d_dst.ptr, d_idx.ptr, d_src.ptr, n, bins, TYPE[np.dtype(dtype)]), "scatter")
    assert np.array_equal(d_dst.get(), tf_oracle.scatter_add(dst0, idx, src))


# ---- matmul --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (128, 128, 16), (129, 131, 17), (300, 200, 1000), (48, 12, 128)])
def test_matmul_ffma(tfcuda_lib, m, n, k):
    rng = np.random.default_rng(m * n + k)
    a, b = rng.random((m, k), dtype=np.float32), rng.random((k, n), dtype=np.float32)
    d_a, d_b, d_c = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.zeros((m, n), np.float32))
    abi.check(tfcuda_lib.tfcuda_matmul(d_a.ptr, d_b.ptr, d_c.ptr, 1, m, n, k, 2), "matmul")
    # same products in the same k order as the reference; FMA keeps the product unrounded -> 1e-6, far inside 1e-3
    assert rel_err(d_c.get(), tf_oracle.matmul(a, b)) <= 1e-6


@pytest.mark.parametrize("m,n,k", [(128, 32, 32), (128, 256, 64), (1000, 200, 48), (4096, 128, 48), (4096, 12, 128), (1024, 1024, 1024), (300, 260, 1000), (77, 52, 36)])
@pytest.mark.parametrize("mode", [0, 1])
def test_matmul_tcgen05(tfcuda_lib, m, n, k, mode):
    """tcgen05 kind::tf32 path.  mode 0 (single TF32 product): operands keep 10 mantissa bits, the hardware truncates ->
    1e-3 class (north_star's matmul bar); mode 1 (3xTF32 split) restores fp32-level accuracy: 5e-5 here (the tensor core's
    accumulator adds round toward zero, which shows at K ~ 1000)."""
    rng = np.random.default_rng(m + n + k)
    for dist in ("uniform", "normal"):
        a = rng.random((m, k), dtype=np.float32) if dist == "uniform" else rng.standard_normal((m, k)).astype(np.float32)
        b = rng.random((k, n), dtype=np.float32) if dist == "uniform" else rng.standard_normal((k, n)).astype(np.float32)
        d_a, d_b, d_c = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.full((m, n), np.nan, np.float32))
        abi.check(tfcuda_lib.tfcuda_matmul(d_a.ptr, d_b.ptr, d_c.ptr, 1, m, n, k, mode), "matmul")
        want = a.astype(np.float64) @ b.astype(np.float64)
        assert rel_err(d_c.get(), want) <= (1e-3 if mode == 0 else 5e-5)


def test_matmul_3xtf32_inf_and_large_magnitudes(tfcuda_lib):
    """mode 1 splits x = hi + lo: an Inf input must give Inf (as the reference's fp32 loop), not NaN rows from lo = Inf - Inf; large
    finite magnitudes keep the fp32-level accuracy."""
    rng = np.random.default_rng(4)
    m, n, k = 128, 64, 64
    a, b = rng.random((m, k), dtype=np.float32), rng.random((k, n), dtype=np.float32)
    a[3, 5] = np.inf
    a[7] *= np.float32(1e30)
    d_a, d_b, d_c = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.zeros((m, n), np.float32))
    abi.check(tfcuda_lib.tfcuda_matmul(d_a.ptr, d_b.ptr, d_c.ptr, 1, m, n, k, 1), "matmul")
    got = d_c.get()
    with np.errstate(over="ignore", invalid="ignore"):
        want = a.astype(np.float64) @ b.astype(np.float64)
    assert np.all(np.isposinf(got[3])) and not np.isnan(got).any()
    rows = [r for r in range(m) if r not in (3, 7)]
    assert rel_err(got[rows], want[rows]) <= 5e-5 and rel_err(got[[7]], want[[7]]) <= 5e-5


def test_matmul_tcgen05_batched_and_unaligned_fallback(tfcuda_lib):
    rng = np.random.default_rng(9)
    a, b = rng.random((3, 130, 64), dtype=np.float32), rng.random((3, 64, 96), dtype=np.float32)
    d_a, d_b, d_c = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.zeros((3, 130, 96), np.float32))
    abi.check(tfcuda_lib.tfcuda_matmul(d_a.ptr, d_b.ptr, d_c.ptr, 3, 130, 96, 64, 1), "matmul")
    assert rel_err(d_c.get(), a.astype(np.float64) @ b.astype(np.float64)) <= 5e-5
    # K = 30: the row pitch is not a multiple of 16 bytes -> TMA cannot describe it -> FFMA kernel (still our CUDA path)
    a, b = rng.random((50, 30), dtype=np.float32), rng.random((30, 20), dtype=np.float32)
    d_a, d_b, d_c = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.zeros((50, 20), np.float32))
    abi.check(tfcuda_lib.tfcuda_matmul(d_a.ptr, d_b.ptr, d_c.ptr, 1, 50, 20, 30, 0), "matmul")
    assert rel_err(d_c.get(), tf_oracle.matmul(a, b)) <= 1e-6


@pytest.mark.parametrize("r,m,n", [(1, 1, 1), (5, 3, 2), (4097, 48, 128), (10000, 128, 12), (3000, 130, 70), (777, 20, 24), (100000, 48, 128), (65536, 128, 12)])
def test_matmul_tn(tfcuda_lib, r, m, n):
    """C = A^T @ B contracted over the leading extent (the weight-gradient kernel).  fp32 FFMA products, split-K partial sums added
    in a fixed order: 2e-6 of the result scale against float64 (far inside north_star's 1e-3 matmul/gradient bar), and bit-identical
    between two runs (no float atomics)."""
    rng = np.random.default_rng(r + m + n)
    a, b = rng.standard_normal((r, m)).astype(np.float32), rng.standard_normal((r, n)).astype(np.float32)
    d_a, d_b = abi.DeviceArray(a), abi.DeviceArray(b)
    d_c, d_c2 = abi.DeviceArray(np.full((m, n), np.nan, np.float32)), abi.DeviceArray(np.full((m, n), np.nan, np.float32))
    abi.check(tfcuda_lib.tfcuda_matmul_tn(d_a.ptr, d_b.ptr, d_c.ptr, r, m, n), "matmul_tn")
    abi.check(tfcuda_lib.tfcuda_matmul_tn(d_a.ptr, d_b.ptr, d_c2.ptr, r, m, n), "matmul_tn")
    want = a.astype(np.float64).T @ b.astype(np.float64)
    got = d_c.get()
    assert rel_err(got, want) <= 2e-6
    assert np.array_equal(got.view(np.uint32), d_c2.get().view(np.uint32))
    assert rel_err(got, tf_oracle.matmul(np.ascontiguousarray(a.T), b)) <= 1e-5 if r <= 10000 else True


def test_matmul_tn_large_output_extent(tfcuda_lib):
    """M > 32767 (an embedding-sized weight): outside the split-K kernel's range, served by transpose + dense matmul (fp32-accurate mode)."""
    r, m, n = 96, 40000, 8
    rng = np.random.default_rng(9)
    a, b = rng.standard_normal((r, m)).astype(np.float32), rng.standard_normal((r, n)).astype(np.float32)
    d_a, d_b, d_c = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.full((m, n), np.nan, np.float32))
    abi.check(tfcuda_lib.tfcuda_matmul_tn(d_a.ptr, d_b.ptr, d_c.ptr, r, m, n), "matmul_tn (large M)")
    want = a.astype(np.float64).T @ b.astype(np.float64)
    assert rel_err(d_c.get(), want) <= 1e-5


@pytest.mark.parametrize("r,k,n", [(1, 1, 1), (70, 48, 128), (5000, 48, 128), (5000, 128, 12), (5000, 12, 128), (5000, 128, 48), (257, 7, 5), (4097, 36, 30), (100000, 128, 128)])
def test_matmul_rows(tfcuda_lib, r, k, n):
    """Skinny matmul with the weight matrix resident in shared memory: fp32 FFMA in the reference's k order -> 1e-6 of the result scale
    against the restatement (FMA keeps the product unrounded), like test_matmul_ffma."""
    assert tfcuda_lib.tfcuda_matmul_rows_supported(r, k, n)
    rng = np.random.default_rng(r + k + n)
    a, b = rng.random((r, k), dtype=np.float32), rng.random((k, n), dtype=np.float32)
    d_a, d_b, d_c = abi.DeviceArray(a), abi.DeviceArray(b), abi.DeviceArray(np.full((r, n), np.nan, np.float32))
    abi.check(tfcuda_lib.tfcuda_matmul_rows(d_a.ptr, d_b.ptr, d_c.ptr, r, k, n), "matmul_rows")
    want = tf_oracle.matmul(a, b) if r <= 5000 else (a.astype(np.float64) @ b.astype(np.float64))
    assert rel_err(d_c.get(), want) <= (1e-6 if r <= 5000 else 2e-6)


# ---- n-body --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 255, 256, 257, 1500])
def test_nbody_step(tfcuda_lib, n):
    """n < 2048: the scalar kernel (the packed f32x2 kernel, n >= 2048, is tested in tests/test_zy_late_gpu.py)."""
    rng = np.random.default_rng(n)
    x = (5.0 * rng.standard_normal((n, 3))).astype(np.float32)
    v = (0.1 * rng.standard_normal((n, 3))).astype(np.float32)
    d = [abi.DeviceArray(t) for t in (x, v, np.zeros_like(x), np.zeros_like(v))]
    abi.check(tfcuda_lib.tfcuda_nbody_step(d[0].ptr, d[1].ptr, d[2].ptr, d[3].ptr, n, 0.001, 1e-4), "nbody")
    xn, vn = tf_oracle.nbody_step(x, v)
    assert rel_err(d[3].get(), vn) <= 1e-5 and rel_err(d[2].get(), xn) <= 1e-6
