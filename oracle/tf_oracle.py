"""CPU restatement (numpy) of the reference algorithms that the hand-written library kernels replace.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (tensorfrost_b200/, the CUDA module, libtfcuda.so) imports or
calls this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may.

Parity pin: every function below is checked in tests/test_oracle.py against golden outputs produced by the REFERENCE
ITSELF (its C++/OpenMP backend, built from /root/reference by oracle/build_ref.sh and run by
tests/golden/make_golden.py).  The reference ships no golden vectors of its own, so those fixtures
are the pin; the second, stronger oracle is the reference module in oracle/_ref, run live by the GPU tests.

Each function cites the reference code it restates (paths relative to the reference root).
"""
import numpy as np

f32 = np.float32


# ---------------------------------------------------------------------------------------------------------------
# radix sort — Python/TensorFrost/sort.py:38-187
# ---------------------------------------------------------------------------------------------------------------
def map_key_to_uint(keys):
    """sort.py:52-72: order-preserving bijections onto uint32 (float: flip all bits of negatives, the sign bit of
    non-negatives; int: flip the sign bit)."""
    keys = np.asarray(keys)
    if keys.dtype == np.float32:
        u = keys.view(np.uint32)
        mask = np.where((u >> 31) == 1, np.uint32(0xFFFFFFFF), np.uint32(0x80000000))
        return u ^ mask
    if keys.dtype == np.int32:
        return keys.view(np.uint32) ^ np.uint32(0x80000000)
    return keys.astype(np.uint32, copy=False)


def map_uint_to_key(u, dtype):
    if dtype == np.float32:
        mask = np.where((u >> 31) == 0, np.uint32(0xFFFFFFFF), np.uint32(0x80000000))
        return (u ^ mask).view(np.float32)
    if dtype == np.int32:
        return (u ^ np.uint32(0x80000000)).view(np.int32)
    return u


def radix_sort(keys, values=None, bits_per_pass=6, max_bits=32):
    """LSD radix sort, one stable counting pass per digit (sort.py:89-176: histogram -> exclusive scan -> each element
    goes to offset[digit] + number of earlier elements with the same digit).  The reference always runs an EVEN number
    of passes (`tf.loop(iters // 2)` around two ping-pong iterations, sort.py:102,174-175), restated here."""
    keys = np.asarray(keys)
    dtype = keys.dtype
    u = map_key_to_uint(keys).copy()
    vals = None if values is None else np.asarray(values).copy()
    iters = (max_bits + bits_per_pass - 1) // bits_per_pass
    radix = 1 << bits_per_pass
    for it in range(2 * (iters // 2)):
        shift = it * bits_per_pass
        digit = ((u >> np.uint32(shift)) if shift < 32 else np.zeros_like(u)) & np.uint32(radix - 1)
        # rank among equal digits in input order == stable counting sort
        order = np.argsort(digit, kind="stable")
        dest = np.empty_like(order)
        dest[order] = np.arange(len(u))
        # (dest[i] == start[digit[i]] + #{j < i : digit[j] == digit[i]}, the offset sort.py:146-156 computes per element)
        out = np.empty_like(u)
        out[dest] = u
        u = out
        if vals is not None:
            vout = np.empty_like(vals)
            vout[dest] = vals
            vals = vout
    k = map_uint_to_key(u, dtype)
    return (k, vals) if vals is not None else k


# ---------------------------------------------------------------------------------------------------------------
# reductions — Compiler/Implementations.cpp:243-303 (serial accumulate in index order) and :360-440 (ops)
# ---------------------------------------------------------------------------------------------------------------
def reduce(a, axis=-1, op="sum", staged_chunk=None):
    """Per output element: acc = initial; for k in range(n): acc = op(acc, a[..., k, ...]) in the element type.
    staged_chunk=128 reproduces the two-stage split the reference applies to constant axes >= 1024
    (Steps/Optimization.cpp:469-510): partial results over chunks of 128, then a reduction of the partials."""
    a = np.asarray(a)
    axis = axis % a.ndim
    n = a.shape[axis]
    moved = np.moveaxis(a, axis, 0)
    if staged_chunk and n >= 1024 and n % staged_chunk == 0 and op in ("sum", "max", "min", "mean", "norm"):
        # SplitDim(input, 128, axis) -> index = g*128 + e; the FIRST stage reduces each contiguous chunk of 128
        # (the internal axis order is reversed, so `axis` is the e dimension), the second reduces the n/128 partials
        # with the SAME op (so mean = mean of chunk means) (Optimization.cpp:496-501)
        parts = moved.reshape((n // staged_chunk, staged_chunk) + moved.shape[1:])
        if op == "norm":
            return np.sqrt(reduce(reduce(parts * parts, axis=1, op="sum"), axis=0, op="sum")).astype(np.float32)
        return reduce(reduce(parts, axis=1, op=op), axis=0, op=op)
    if op in ("sum", "mean", "norm"):
        acc = np.zeros(moved.shape[1:], dtype=a.dtype)
        src = moved * moved if op == "norm" else moved
        for k in range(n):
            acc = (acc + src[k]).astype(a.dtype)
        if op == "mean":
            return (acc / f32(n)).astype(np.float32)
        if op == "norm":
            return np.sqrt(acc).astype(np.float32)
        return acc
    if op == "max":
        init = {np.dtype(np.float32): -np.finfo(np.float32).max, np.dtype(np.int32): np.iinfo(np.int32).min}.get(a.dtype, 0)
        acc = np.full(moved.shape[1:], init, dtype=a.dtype)
        for k in range(n):
            acc = np.where(acc > moved[k], acc, moved[k])  # max(a,b) = a > b ? a : b  (CPP.cpp:38-51)
        return acc
    if op == "min":
        init = {np.dtype(np.float32): np.finfo(np.float32).max, np.dtype(np.int32): np.iinfo(np.int32).max}.get(a.dtype, 0xFFFFFFFF)
        acc = np.full(moved.shape[1:], init, dtype=a.dtype)
        for k in range(n):
            acc = np.where(acc < moved[k], acc, moved[k])
        return acc
    if op == "any":
        return (moved != 0).any(axis=0).astype(a.dtype)
    if op == "all":
        return (moved != 0).all(axis=0).astype(a.dtype)
    raise ValueError(op)


# ---------------------------------------------------------------------------------------------------------------
# scan — Compiler/Implementations.cpp:305-359
# ---------------------------------------------------------------------------------------------------------------
def prefix_sum(a, axis=-1):
    a = np.asarray(a)
    axis = axis % a.ndim
    moved = np.moveaxis(a, axis, 0)
    out = np.empty_like(moved)
    acc = np.zeros(moved.shape[1:], dtype=a.dtype)
    for k in range(moved.shape[0]):
        acc = (acc + moved[k]).astype(a.dtype)
        out[k] = acc
    return np.moveaxis(out, 0, axis)


# ---------------------------------------------------------------------------------------------------------------
# matmul — Compiler/Implementations.cpp:560-646: c = 0; for k: c = c + a[i,k]*b[k,j]  (fp32 multiply, fp32 add)
# ---------------------------------------------------------------------------------------------------------------
def matmul(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    c = np.zeros(a.shape[:-1] + (b.shape[-1],), dtype=np.float32)
    for k in range(a.shape[-1]):
        c = (c + (a[..., :, k:k + 1] * b[..., k:k + 1, :]).astype(np.float32)).astype(np.float32)
    return c


# ---------------------------------------------------------------------------------------------------------------
# weight gradient of C = A @ B for a 2-D B and an N-D A — the matmul VJP, Compiler/Implementations.cpp:133-135:
#   dB = Matmul(Transpose(A), dC) evaluated per leading slice (ComputeMatMul :560-646, serial k loop), then summed over the
#   leading axes one at a time, outermost first (ReduceGradientToShape :5-54 -> ComputeReduction :243-303, serial loops).
# ---------------------------------------------------------------------------------------------------------------
def matmul_weight_grad(a, dc):
    a = np.asarray(a, dtype=np.float32)
    dc = np.asarray(dc, dtype=np.float32)
    per_slice = matmul(np.swapaxes(a, -1, -2), dc)  # [batch..., K, N]
    while per_slice.ndim > 2:
        acc = np.zeros(per_slice.shape[1:], dtype=np.float32)
        for i in range(per_slice.shape[0]):
            acc = (acc + per_slice[i]).astype(np.float32)
        per_slice = acc
    return per_slice


# ---------------------------------------------------------------------------------------------------------------
# scatter-add — tf.scatterAdd / InterlockedAdd (CPP.cpp:141-158); indices clamp (Steps/GraphOps.cpp:1022-1025)
# ---------------------------------------------------------------------------------------------------------------
def scatter_add(dst, index, src):
    dst = np.array(dst, copy=True)
    idx = np.clip(np.asarray(index, dtype=np.int64), 0, dst.size - 1)
    flat = dst.reshape(-1)
    np.add.at(flat, idx, np.asarray(src, dtype=dst.dtype))
    return dst


# ---------------------------------------------------------------------------------------------------------------
# n-body — examples/Simulation/n-body-benchmark.py:16-34
# ---------------------------------------------------------------------------------------------------------------
def nbody_step(x, v, dt=0.001, eps=1e-4):
    x = np.asarray(x, dtype=np.float32)
    v = np.asarray(v, dtype=np.float32)
    n = x.shape[0]
    force = np.zeros_like(x)
    for j in range(n):  # sum over axis=1 in j order, fp32
        dx = x - x[j]
        d2 = ((dx[:, 0] * dx[:, 0] + dx[:, 1] * dx[:, 1]).astype(f32) + dx[:, 2] * dx[:, 2]).astype(f32) + f32(eps)
        dist = np.sqrt(d2).astype(f32)
        force = (force + ((-dx * f32(1.0)) / (d2 * dist)[:, None]).astype(f32)).astype(f32)
    v_new = (v + force * f32(dt)).astype(f32)
    x_new = (x + v_new * f32(dt)).astype(f32)
    return x_new, v_new
