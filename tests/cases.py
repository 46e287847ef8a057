"""Named TensorFrost programs ("cases") that the parity tests run on BOTH backends.

Every case is backend-agnostic python written against the public `tf.*` API; the same source is traced
by the oracle module (`tf.cpu`, the reference's C++/OpenMP backend) and by the CUDA module (`tf.cuda`).
The cases follow the reference's own test programs (tests/sorting_test.py, linalg_test.py,
reshape_reduction_test.py, split_dim_test.py, autograd_test.py) and add what those leave unpinned
(SURVEY.md §4 "gaps"): value stability of sorts, scans, atomics, math ops, control flow, host read/write.

A case is  Case(name, build(tf) -> program, make_inputs(rng, size) -> [np arrays], kind, tol)  with
kind "exact" (bit-exact: integer / index / sort / scan work) or "float" (relative tolerance `tol`).
Inputs are generated from np.random.default_rng(seed) only, so fixtures are reproducible from (name, seed, size).
"""
from dataclasses import dataclass
from typing import Callable, List, Optional

import numpy as np


@dataclass
class Case:
    name: str
    build: Callable          # build(tf) -> compiled program (or a callable taking *tensors)
    make_inputs: Callable    # make_inputs(rng, size) -> list of numpy arrays
    kind: str = "float"      # "exact" | "float"
    tol: float = 1e-5        # element-wise relative bar for kind == "float": |got - ref| <= tol * max(|ref|, floor * max|ref|)
    floor: float = 0.05      # absolute floor as a fraction of the output's max magnitude: elements below it are sums / differences
                             # of O(max) terms (stencils, residuals, gradients), their error scales with the terms, not with them
    default_size: int = 64
    outputs: Optional[List[str]] = None


CASES = {}


def case(name, kind="float", tol=1e-5, default_size=64, outputs=None, floor=0.05):
    def deco(fn):
        build, make_inputs = fn()
        CASES[name] = Case(name, build, make_inputs, kind, tol, floor, default_size, outputs)
        return fn
    return deco


# ---------------------------------------------------------------------------------------------
# elementwise / stencil
# ---------------------------------------------------------------------------------------------
@case("wave", tol=1e-6, default_size=96)
def _wave():
    def build(tf):
        def step():
            u = tf.input([-1, -1], tf.float32)
            v = tf.input(u.shape, tf.float32)
            i, j = u.indices
            lap = u[i - 1, j] + u[i + 1, j] + u[i, j - 1] + u[i, j + 1] - u * 4.0
            v_new = v + lap * 0.2
            return u + v_new * 0.2, v_new
        return tf.compile(step)

    def make_inputs(rng, size):
        return [rng.random((size, size + 7), dtype=np.float32), rng.random((size, size + 7), dtype=np.float32)]
    return build, make_inputs


@case("math_ops", tol=2e-6, default_size=4096)
def _math_ops():
    """Every float function op of the op table (Operations.cpp:120-175) that the oracle can compile."""
    def build(tf):
        def prog():
            x = tf.input([-1], tf.float32)      # in (0.05, 4)
            y = tf.input(x.shape, tf.float32)   # in (-1, 1)
            outs = [
                tf.exp(y), tf.exp2(y), tf.log(x), tf.log2(x), tf.sqrt(x), tf.sin(x), tf.cos(x), tf.tan(y),
                tf.asin(y), tf.acos(y), tf.atan(x), tf.sinh(y), tf.cosh(y), tf.tanh(x), tf.pow(x, y), tf.atan2(y, x),
                tf.abs(y), tf.sign(y), tf.ceil(x * 3.0), tf.floor(y * 3.0), tf.round(y * 3.0),
                tf.min(x, y), tf.max(x, y), tf.clamp(y, -0.25, 0.5), tf.lerp(x, y, 0.3), tf.fma(x, y, x),
                tf.smoothstep(0.0, 1.0, y), tf.select(y > 0.0, x, -x), x / (y * y + 0.5), -x + y * x - 2.0,
                tf.float(tf.int(x * 10.0)), tf.float(tf.uint(x * 7.0)), tf.float(y > 0.1),
            ]
            return outs
        return tf.compile(prog)

    def make_inputs(rng, size):
        x = (rng.random(size, dtype=np.float32) * 3.95 + 0.05).astype(np.float32)
        y = (rng.random(size, dtype=np.float32) * 1.98 - 0.99).astype(np.float32)
        return [x, y]
    return build, make_inputs


@case("int_ops", kind="exact", default_size=4096)
def _int_ops():
    """Integer / bit / hash ops: bit-exact class."""
    def build(tf):
        def prog():
            a = tf.input([-1], tf.int32)
            b = tf.input(a.shape, tf.int32)     # in [1, 31]
            u = tf.input(a.shape, tf.uint32)
            outs = [
                a + b, a - b, a * b, a / b, a % b, a & b, a | b, a ^ b, ~a, -a, a << (b % 8), a >> (b % 8),
                u + tf.uint(b), u * u, u / tf.uint(b), u % tf.uint(b), u >> tf.uint(b % 16), u << tf.uint(b % 16), u ^ (u >> 7),
                tf.pcg(u), tf.reversebits(u), tf.min(a, b), tf.max(a, b), tf.abs(a), tf.sign(a), tf.clamp(a, -100, 100),
                tf.int(a < b), tf.int((a < b) & (a > -b)), tf.int((a == b) | (a >= 0)), tf.select(a != b, a, b),
                tf.asint(tf.asuint(tf.asfloat(u))), tf.int(u), tf.uint(a), tf.int(tf.float(a % 1000)),
                # (asint(float) is a VALUE conversion on the oracle - its helper header only has asint(uint) - so unpinned)
            ]
            return outs
        return tf.compile(prog)

    def make_inputs(rng, size):
        a = rng.integers(-2 ** 20, 2 ** 20, size, dtype=np.int32)
        b = rng.integers(1, 32, size, dtype=np.int32)
        u = rng.integers(0, 2 ** 32, size, dtype=np.uint64).astype(np.uint32)
        return [a, b, u]
    return build, make_inputs


@case("pcgf_random", tol=0.0, default_size=8192)
def _pcgf():
    """tf.pcgf: float(pcg(v)) / float(0xffffffff) (CPP.cpp:268-271); exact arithmetic on both sides."""
    def build(tf):
        def prog():
            u = tf.input([-1], tf.uint32)
            return tf.pcgf(u), tf.pcgf(tf.pcg(u) + tf.uint(u.indices[0]))
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.integers(0, 2 ** 32, size, dtype=np.uint64).astype(np.uint32)]
    return build, make_inputs


@case("control_flow", kind="exact", default_size=2048)
def _control_flow():
    """In-kernel loop / if / break / continue and a per-thread local buffer (Collatz-style integer work)
    inside an explicit tf.kernel scope."""
    def build(tf):
        def prog():
            seed = tf.input([-1], tf.int32)
            out_steps = tf.buffer(seed.shape, tf.int32)
            out_acc = tf.buffer(seed.shape, tf.int32)
            out_last = tf.buffer(seed.shape, tf.int32)
            out_hist = tf.buffer(seed.shape, tf.int32)
            with tf.kernel(seed.shape) as i:
                n = tf.const(0)
                n.val = seed[i]
                steps = tf.const(0)
                acc = tf.const(0)
                hist = tf.local_buffer(4, tf.int32)
                for k in range(4):
                    hist[k] = 0
                with tf.loop(200) as it:
                    with tf.if_cond(n == 1):
                        tf.break_loop()
                    steps.val += 1
                    with tf.if_cond((n % 2) == 0):
                        n.val = n / 2
                        tf.continue_loop()
                    n.val = n * 3 + 1
                    acc.val += it
                    hist[it % 4] = hist[it % 4] + 1
                out_steps[i] = steps
                out_acc[i] = acc
                out_last[i] = n
                out_hist[i] = hist[0] + 2 * hist[1] + 3 * hist[2] + 5 * hist[3]
            return out_steps, out_acc, out_last, out_hist
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.integers(1, 5000, size, dtype=np.int32)]
    return build, make_inputs


# ---------------------------------------------------------------------------------------------
# reductions / scans / reshapes
# ---------------------------------------------------------------------------------------------
@case("reshape_reduce", tol=2e-6, default_size=10)
def _reshape_reduce():
    """reshape + last-axis max/min/sum/mean/norm, chained reductions, mean over axis 0
    (the reference's tests/reshape_reduction_test.py:7-21)."""
    def build(tf):
        def prog():
            a = tf.input([-1, -1, -1, -1], tf.float32)
            n, bx, by, bz = a.shape
            flat = tf.reshape(a, [n, bx * by * bz])
            mx, mn = tf.max(flat), tf.min(flat)
            return [mx, mn, tf.sum(flat), tf.mean(flat), tf.norm(flat), tf.max(mx), tf.min(mn), tf.mean(a, axis=0), tf.sum(a, axis=2)]
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.random((size, 5, 6, 7), dtype=np.float32)]
    return build, make_inputs


@case("row_reductions", tol=1e-5, default_size=2048)
def _row_reductions():
    """C4b: sum / max / mean / norm over the last axis of a square fp32 matrix (constant axis >= 1024 takes the
    oracle's staged path, Steps/Optimization.cpp:469-510)."""
    def build(tf):
        def make(n):
            def prog():
                a = tf.input([-1, n], tf.float32)
                return tf.sum(a), tf.max(a), tf.mean(a), tf.norm(a)
            return tf.compile(prog)
        cache = {}

        def run(a):
            n = a.shape[1]
            if n not in cache:
                cache[n] = make(n)
            return cache[n](a)
        return run

    def make_inputs(rng, size):
        return [rng.random((size // 2, size), dtype=np.float32)]
    return build, make_inputs


@case("int_reductions", kind="exact", default_size=300)
def _int_reductions():
    def build(tf):
        def prog():
            a = tf.input([-1, -1], tf.int32)
            u = tf.input(a.shape, tf.uint32)
            # (uint max/min do not compile on the oracle: min(uint,uint) is ambiguous in its helper header -> unpinned)
            return tf.sum(a), tf.max(a), tf.min(a), tf.sum(a, axis=0), tf.sum(u), tf.sum(u, axis=0)
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.integers(-1000, 1000, (size, size + 3), dtype=np.int32),
                rng.integers(0, 2 ** 20, (size, size + 3), dtype=np.uint64).astype(np.uint32)]
    return build, make_inputs


@case("prefix_sum", kind="exact", default_size=3000)
def _prefix_sum():
    """tf.prefix_sum (np.cumsum) on ints along both axes and on integer-valued floats (exact in fp32)."""
    def build(tf):
        def prog():
            a = tf.input([-1, -1], tf.int32)
            f = tf.input([-1], tf.float32)
            return tf.prefix_sum(a), tf.prefix_sum(a, axis=0), tf.prefix_sum(f)
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.integers(-50, 50, (7, size), dtype=np.int32), rng.integers(0, 8, size * 5).astype(np.float32)]
    return build, make_inputs


@case("split_merge", kind="exact", default_size=128)
def _split_merge():
    """split_dim / merge_dim on int32 (tests/split_dim_test.py:7-11)."""
    def build(tf):
        def prog():
            data = tf.input([-1, -1, -1], tf.int32)
            parts = tf.split_dim(data, 32, 0)
            return tf.merge_dim(parts, axis=1), parts
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.integers(0, 100, (size, 24, 8), dtype=np.int32)]
    return build, make_inputs


# ---------------------------------------------------------------------------------------------
# sorting (Python/TensorFrost/sort.py)
# ---------------------------------------------------------------------------------------------
def _sort_case(name, key_dtype, algo):
    @case(name, kind="exact", default_size=1 << 14)
    def _c():
        def build(tf):
            tft = {np.uint32: tf.uint32, np.float32: tf.float32, np.int32: tf.int32}[key_dtype]

            def prog():
                keys = tf.input([-1], tft)
                values = tf.input([-1], tf.uint32)
                fn = tf.sort.radix if algo == "radix" else tf.sort.bitonic
                k, v = fn(keys, values)
                return k, v
            return tf.compile(prog)

        def make_inputs(rng, size):
            if key_dtype is np.uint32:
                keys = rng.integers(0, 2 ** 32, size, dtype=np.uint64).astype(np.uint32)
                keys[: size // 8] = keys[size // 8: 2 * (size // 8)]  # duplicates: exercises stability
            elif key_dtype is np.int32:
                keys = rng.integers(-2 ** 31, 2 ** 31, size, dtype=np.int64).astype(np.int32)
                keys[: size // 8] = keys[size // 8: 2 * (size // 8)]
            else:
                keys = (rng.standard_normal(size) * 1e3).astype(np.float32)
                keys[: size // 8] = keys[size // 8: 2 * (size // 8)]
                keys[-3:] = [0.0, -0.0, np.float32(np.inf)]
            return [keys, np.arange(size, dtype=np.uint32)]
        return build, make_inputs
    return _c


_sort_case("sort_radix_u32", np.uint32, "radix")
_sort_case("sort_radix_f32", np.float32, "radix")
_sort_case("sort_radix_i32", np.int32, "radix")
_sort_case("sort_bitonic_u32", np.uint32, "bitonic")


# ---------------------------------------------------------------------------------------------
# atomics (tf.scatter*)
# ---------------------------------------------------------------------------------------------
@case("atomics", kind="exact", default_size=20000)
def _atomics():
    """All scatter ops on int/uint/float destinations with heavy collisions.  Float adds use integer-valued
    floats so the result does not depend on the (non-deterministic) order of the atomics."""
    def build(tf):
        def prog():
            idx = tf.input([-1], tf.int32)       # in [0, 64)
            vi = tf.input(idx.shape, tf.int32)
            vf = tf.input(idx.shape, tf.float32)
            vu = tf.input(idx.shape, tf.uint32)
            e, = idx.indices
            add_i, min_i, max_i = tf.zeros([64], tf.int32), tf.zeros([64], tf.int32), tf.zeros([64], tf.int32)
            add_f, min_f, max_f = tf.zeros([64], tf.float32), tf.zeros([64], tf.float32), tf.zeros([64], tf.float32)
            add_u, or_u, xor_u, and_u = tf.zeros([64], tf.uint32), tf.zeros([64], tf.uint32), tf.zeros([64], tf.uint32), tf.zeros([64], tf.uint32)
            b, = and_u.indices
            and_u[b] = ~tf.uint(0)
            tf.scatterAdd(add_i[idx[e]], vi[e])
            tf.scatterMin(min_i[idx[e]], vi[e])
            tf.scatterMax(max_i[idx[e]], vi[e])
            tf.scatterAdd(add_f[idx[e]], vf[e])
            tf.scatterMin(min_f[idx[e]], vf[e])
            tf.scatterMax(max_f[idx[e]], vf[e])
            tf.scatterAdd(add_u[idx[e]], vu[e])
            tf.scatterOr(or_u[idx[e]], vu[e])
            tf.scatterXor(xor_u[idx[e]], vu[e])
            tf.scatterAnd(and_u[idx[e]], vu[e] | tf.uint(0xffff0000))
            return add_i, min_i, max_i, add_f, min_f, max_f, add_u, or_u, xor_u, and_u
        return tf.compile(prog)

    def make_inputs(rng, size):
        idx = rng.integers(0, 64, size, dtype=np.int32)
        idx[: size // 4] = 3  # one hot address
        vi = rng.integers(-100, 100, size, dtype=np.int32)
        vf = rng.integers(-8, 9, size).astype(np.float32)
        vu = rng.integers(0, 2 ** 32, size, dtype=np.uint64).astype(np.uint32)
        return [idx, vi, vf, vu]
    return build, make_inputs


@case("scatter_matmul", tol=2e-5, default_size=48)
def _scatter_matmul():
    """examples/Algorithms/scatter.py idea: C[i,j] += A[i,k]*B[k,j] through float atomics (order dependent -> tolerance)."""
    def build(tf):
        def prog():
            a = tf.input([-1, -1], tf.float32)
            n, m = a.shape
            b = tf.input([m, -1], tf.float32)
            c = tf.zeros([n, b.shape[1]])
            i, j, k = tf.indices([n, b.shape[1], m])
            tf.scatterAdd(c[i, j], a[i, k] * b[k, j])
            return c
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.random((size, size + 5), dtype=np.float32), rng.random((size + 5, size - 3), dtype=np.float32)]
    return build, make_inputs


# ---------------------------------------------------------------------------------------------
# linear algebra
# ---------------------------------------------------------------------------------------------
@case("matmul", tol=1e-5, default_size=160)
def _matmul():
    def build(tf):
        def prog():
            a = tf.input([-1, -1], tf.float32)
            b = tf.input([a.shape[1], -1], tf.float32)
            return a @ b, tf.transpose(a), tf.dot(a, a)
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.random((size + 9, size), dtype=np.float32), rng.random((size, size - 17), dtype=np.float32)]
    return build, make_inputs


@case("qr_inverse", tol=2e-4, default_size=5)
def _qr_inverse():
    """Gram-Schmidt QR + triangular inverse + matmul with host loops around kernels
    (the reference's tests/linalg_test.py:39-90)."""
    def build(tf):
        def gram_schmidt(a):
            m, n = a.shape
            q, r = tf.zeros([m, n]), tf.zeros([n, n])
            row = tf.index(0, [m])
            with tf.loop(n - 1) as c:
                r[c, c] = tf.norm(a[row, c])
                q[row, c] = a[row, c] / r[c, c]
                p, k = tf.index_grid([0, c + 1], [m, n])
                proj = tf.sum(q[p, c] * a[p, k], axis=0)
                r[c, proj.indices[0] + c + 1] = proj
                a[p, k] -= q[p, c] * r[c, k]
            r[n - 1, n - 1] = tf.norm(a[row, n - 1])
            q[row, n - 1] = a[row, n - 1] / r[n - 1, n - 1]
            return q, r

        def upper_inverse(r):
            n = r.shape[0]
            low = r.T
            inv = tf.zeros([n, n])
            inv[0, 0] = 1.0 / low[0, 0]
            with tf.loop(1, n) as c:
                inv[c, c] = 1.0 / low[c, c]
                p, k = tf.indices([c, c])
                t, = tf.indices([c])
                inv[c, t] = -tf.sum(low[c, p] * inv[p, k], axis=0) / low[c, c]
            return inv.T

        def prog():
            a = tf.input([-1, -1], tf.float32)
            q, r = gram_schmidt(a)
            r_inv = upper_inverse(r)
            return q, r, r_inv, r_inv @ q.T
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [(rng.random((size, size), dtype=np.float32) + np.eye(size, dtype=np.float32))]
    return build, make_inputs


# ---------------------------------------------------------------------------------------------
# n-body (examples/Simulation/n-body-benchmark.py:16-65), both formulations
# ---------------------------------------------------------------------------------------------
@case("nbody", tol=1e-5, default_size=512)
def _nbody():
    def build(tf):
        def prog():
            x = tf.input([-1, 3], tf.float32)
            n = x.shape[0]
            v = tf.input([n, 3], tf.float32)
            dx = tf.unsqueeze(x, axis=1) - tf.unsqueeze(x, axis=0)
            d2 = tf.unsqueeze(tf.sum(dx ** 2.0, axis=-1), axis=-1) + 1e-4
            dist = tf.sqrt(d2)
            force = tf.sum(-dx * 1.0 / (d2 * dist), axis=1)
            dt = 0.001
            v_new = v + force * dt
            return x + v_new * dt, v_new
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [(5.0 * rng.standard_normal((size, 3))).astype(np.float32), np.zeros((size, 3), np.float32)]
    return build, make_inputs


@case("nbody_loop", tol=1e-5, default_size=512)
def _nbody_loop():
    def build(tf):
        def prog():
            x = tf.input([-1, 3], tf.float32)
            n = x.shape[0]
            v = tf.input([n, 3], tf.float32)
            f = tf.buffer([n, 3], tf.float32)
            i, = tf.indices([n])
            fx, fy, fz = tf.const(0.0), tf.const(0.0), tf.const(0.0)
            x0, y0, z0 = x[i, 0], x[i, 1], x[i, 2]
            with tf.loop(n) as j:
                dx, dy, dz = x[j, 0] - x0, x[j, 1] - y0, x[j, 2] - z0
                d2 = dx * dx + dy * dy + dz * dz
                g = -dx / (d2 + 1e-4) * 1.0 / tf.sqrt(d2 + 1e-4)
                fx.val += g * dx
                fy.val += g * dy
                fz.val += g * dz
            f[i, 0], f[i, 1], f[i, 2] = fx, fy, fz
            dt = 0.001
            v_new = v + f * dt
            return x + v_new * dt, v_new
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [(5.0 * rng.standard_normal((size, 3))).astype(np.float32), np.zeros((size, 3), np.float32)]
    return build, make_inputs


# ---------------------------------------------------------------------------------------------
# host-side control: loops with kernels inside, tf.read of device scalars
# ---------------------------------------------------------------------------------------------
@case("host_loop", tol=1e-6, default_size=200)
def _host_loop():
    """A host `for` around dispatches whose trip count is READ BACK from a device tensor (tf.read in the host
    program, Generators.cpp:438-461), plus a host-level write."""
    def build(tf):
        def prog():
            field = tf.input([-1, -1], tf.float32)
            count = tf.input([1], tf.int32)
            steps = count[0]
            i, j = field.indices
            cur = tf.buffer(field.shape, tf.float32)
            cur[i, j] = field[i, j]
            nxt = tf.buffer(field.shape, tf.float32)
            with tf.loop(steps):
                # ping-pong: a one-kernel in-place stencil would race between threads
                nxt[i, j] = (cur[i - 1, j] + cur[i + 1, j] + cur[i, j - 1] + cur[i, j + 1]) * 0.25
                cur[i, j] = nxt[i, j] * 0.5 + cur[i, j] * 0.5
            total = tf.sum(tf.sum(cur))
            return cur, total
        return tf.compile(prog)

    def make_inputs(rng, size):
        return [rng.random((size, size // 2), dtype=np.float32), np.array([5], np.int32)]
    return build, make_inputs


# ---------------------------------------------------------------------------------------------
# autodiff (tests/autograd_test.py pattern: gradients exported through a Module)
# ---------------------------------------------------------------------------------------------
@case("autograd_mlp", tol=1e-4, default_size=32)
def _autograd():
    """conv -> max-pool -> GELU -> dense -> log-softmax loss; loss, prediction and tf.grad of every parameter."""
    def build(tf):
        res, ksize, k1, hidden, classes = 12, 3, 4, 16, 10
        r1 = res - ksize + 1
        r1p = r1 // 2

        def gelu(x):
            return 0.5 * x * (1.0 + tf.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * (x * x * x))))

        def log_softmax(x):
            x = x - tf.unsqueeze(tf.max(x))
            return x - tf.log(tf.unsqueeze(tf.sum(tf.exp(x))))

        def conv2d(x, w):
            n, cin, hi, wi = x.shape
            cout, cin, h, ww = w.shape
            b, co, yy, xx, ci, t = tf.indices([n, cout, hi - h + 1, wi - ww + 1, cin, h * ww])
            di, dj = t % ww, t / ww
            return tf.sum(tf.sum(x[b, ci, yy + di, xx + dj] * w[co, ci, di, dj]))

        def max_pool(x):
            b, c, yy, xx, di, dj = tf.indices([x.shape[0], x.shape[1], x.shape[2] / 2, x.shape[3] / 2, 2, 2])
            return tf.max(tf.max(x[b, c, 2 * yy + di, 2 * xx + dj]))

        def prog():
            x = tf.input([-1, res * res], tf.float32)
            n = x.shape[0]
            y = tf.input([n, classes], tf.float32)
            conv_w = tf.input([k1, 1, ksize, ksize], tf.float32)
            fc1 = tf.input([k1 * r1p * r1p, hidden], tf.float32)
            fc1_b = tf.input([hidden], tf.float32)
            fc2 = tf.input([hidden, classes], tf.float32)
            fc2_b = tf.input([classes], tf.float32)
            h = tf.reshape(x, [n, 1, res, res])
            h = gelu(max_pool(conv2d(h, conv_w)))
            h = tf.reshape(h, [n, k1 * r1p * r1p])
            h = gelu(h @ fc1 + fc1_b)
            yhat = h @ fc2 + fc2_b
            loss = tf.mean(tf.sum(-y * log_softmax(yhat)))
            grads = [tf.grad(loss, p) for p in (conv_w, fc1, fc1_b, fc2, fc2_b)]
            return [loss, yhat] + grads
        return tf.compile(prog)

    def make_inputs(rng, size):
        res, ksize, k1, hidden, classes = 12, 3, 4, 16, 10
        r1p = (res - ksize + 1) // 2
        x = rng.standard_normal((size, res * res)).astype(np.float32)
        y = np.eye(classes, dtype=np.float32)[rng.integers(0, classes, size)]
        return [x, y,
                (0.3 * rng.standard_normal((k1, 1, ksize, ksize))).astype(np.float32),
                (0.1 * rng.standard_normal((k1 * r1p * r1p, hidden))).astype(np.float32),
                (0.1 * rng.standard_normal(hidden)).astype(np.float32),
                (0.1 * rng.standard_normal((hidden, classes))).astype(np.float32),
                (0.1 * rng.standard_normal(classes)).astype(np.float32)]
    return build, make_inputs


@case("autograd_batched_dense", tol=1e-4, default_size=6)
def _autograd_batched_dense():
    """Two dense layers applied to every cell of a [n, h, w, c] field (the NCA layer shape, nca.py:50-56) and the gradients of
    both weight matrices and biases: the matmul VJP with an N-D left operand and a 2-D weight (Implementations.cpp:133-135).
    Also a direct `x2.T @ y2` over the flattened rows."""
    def build(tf):
        cin, hidden, cout, h, w = 8, 32, 12, 10, 7

        def prog():
            x = tf.input([-1, h, w, cin], tf.float32)
            n = x.shape[0]
            w1 = tf.input([cin, hidden], tf.float32)
            b1 = tf.input([hidden], tf.float32)
            w2 = tf.input([hidden, cout], tf.float32)
            b2 = tf.input([cout], tf.float32)
            t = tf.input([n, h, w, cout], tf.float32)
            a = x @ w1 + b1
            a = tf.select(a > 0.0, a, 0.01 * a)
            y = a @ w2 + b2
            loss = tf.mean(tf.mean(tf.mean(tf.mean((y - t) ** 2.0))))
            grads = [tf.grad(loss, p) for p in (w1, b1, w2, b2)]
            rows = n * (h * w)
            x2 = tf.reshape(x, [rows, cin])
            t2 = tf.reshape(t, [rows, cout])
            return [loss, y] + grads + [x2.T @ t2]
        return tf.compile(prog)

    def make_inputs(rng, size):
        cin, hidden, cout, h, w = 8, 32, 12, 10, 7
        return [rng.standard_normal((size, h, w, cin)).astype(np.float32),
                (0.3 * rng.standard_normal((cin, hidden))).astype(np.float32), (0.1 * rng.standard_normal(hidden)).astype(np.float32),
                (0.3 * rng.standard_normal((hidden, cout))).astype(np.float32), (0.1 * rng.standard_normal(cout)).astype(np.float32),
                rng.standard_normal((size, h, w, cout)).astype(np.float32)]
    return build, make_inputs


# ---------------------------------------------------------------------------------------------
def run_case(tf, name, seed=0, size=None, program=None):
    """Build (or reuse) the program of a case, run it on seeded inputs, return (outputs as numpy, program)."""
    c = CASES[name]
    rng = np.random.default_rng(seed)
    inputs = c.make_inputs(rng, size or c.default_size)
    prog = program if program is not None else c.build(tf)
    outs = prog(*inputs)
    if not isinstance(outs, (list, tuple)):
        outs = [outs]
    return [np.array(o.numpy) for o in outs], prog


def compare(c: Case, got, want):
    """Raise AssertionError unless `got` matches `want` under the case's bar."""
    assert len(got) == len(want), f"{c.name}: {len(got)} outputs vs {len(want)}"
    for k, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape, f"{c.name}[{k}]: shape {g.shape} vs {w.shape}"
        if c.kind == "exact" or g.dtype.kind in "iub":
            bad = int(np.sum(g.view(np.uint32) != w.view(np.uint32))) if g.dtype.itemsize == 4 else int(np.sum(g != w))
            assert bad == 0, f"{c.name}[{k}]: {bad} of {g.size} words differ (bit-exact class)"
        else:
            gf, wf = g.astype(np.float64), w.astype(np.float64)
            same_special = np.array_equal(np.isnan(gf), np.isnan(wf)) and np.array_equal(np.isinf(gf), np.isinf(wf))
            assert same_special, f"{c.name}[{k}]: NaN/Inf pattern differs"
            fin = np.isfinite(wf)
            scale = max(float(np.max(np.abs(wf[fin]))) if fin.any() else 0.0, 1e-30)
            floor = getattr(c, "floor", 0.05) * scale
            err = float(np.max(np.abs(gf[fin] - wf[fin]) / np.maximum(np.abs(wf[fin]), floor))) if fin.any() else 0.0
            assert err <= c.tol, (f"{c.name}[{k}]: element-wise error {err:.3e} (|got-ref| / max(|ref|, {getattr(c, 'floor', 0.05):g} max|ref|)) "
                                  f"exceeds {c.tol:.1e}")
