"""Ops of the reference's op table (Compiler/Operations.cpp:150-173) that its C++ helper header never defines (CPP.cpp:31-271), so
they do not compile on the oracle: rcp, rsqrt, frac, modf (python API: Frontend/Python/Definitions/TensorFunctions.cpp:27-65) and
step / the `discard` keyword (no python binding: reachable only from generated code).  `tf.trunc` and `tf.sqr` are bound in python but
missing from the op table ("IR Operation not defined" at trace time on every backend, checked below): tf_trunc exists in the prelude
for completeness and is exercised at the generated-code level with step and discard.  "Parity unpinned" by the reference; the CUDA
prelude defines them with their HLSL/GLSL meaning (csrc/prelude.cuh) and this file pins that meaning against numpy on the GPU.
Bar: elementwise |got - want| <= 1e-6 |want| (+ 1e-30), written below."""
import ctypes as C

import numpy as np
import pytest


def _program(tf):
    def prog():
        x = tf.input([-1], tf.float32)      # in (0.05, 4)
        y = tf.input(x.shape, tf.float32)   # in (-3, 3), never 0
        return [tf.rcp(x), tf.rsqrt(x), tf.frac(y * 3.7), tf.modf(y * 5.0, x), tf.rcp(y)]
    return tf.compile(prog)


def _inputs(n=4099):
    rng = np.random.default_rng(11)
    x = (rng.random(n, dtype=np.float32) * 3.95 + 0.05).astype(np.float32)
    y = (rng.random(n, dtype=np.float32) * 6.0 - 3.0).astype(np.float32)
    y[np.abs(y) < 1e-3] = 0.5
    return x, y


def _expected(x, y):
    f = np.float32
    t = (y * f(3.7)).astype(f)
    a = (y * f(5.0)).astype(f)
    return [f(1.0) / x, f(1.0) / np.sqrt(x), t - np.floor(t), a - x * np.floor(a / x), f(1.0) / y]


def _close(got, want, name, tol=1e-6):
    g, w = got.astype(np.float64), want.astype(np.float64)
    err = np.abs(g - w) / (np.abs(w) + 1e-30)
    assert float(err.max()) <= tol, f"{name}: elementwise relative error {err.max():.3e} > {tol:.0e}"


def test_emitted_text_of_unpinned_ops_passes_nvrtc():
    """CPU: the emitter prints tf_rcp / tf_rsqrt / tf_frac / tf_trunc / tf_modf and NVRTC accepts them for sm_100a."""
    import subprocess, sys, os  # noqa: E401
    here = os.path.dirname(os.path.abspath(__file__))
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import tensorfrost_b200, test_op_table_gpu as t\n"
        "from tensorfrost_b200 import abi\n"
        "tf = tensorfrost_b200.import_module(); tf.initialize(tf.codegen, '', tf.cuda_lang)\n"
        "p = t._program(tf)\n"
        "ks = tf.get_all_generated_kernels(); src = '\\n'.join(k[0][1] + k[0][2] for k in ks)\n"
        "assert all(s in src for s in ('tf_rcp', 'tf_rsqrt', 'tf_frac', 'tf_modf')), src\n"
        "for name in ('trunc', 'sqr'):\n"
        "    try:\n"
        "        tf.compile(lambda: [getattr(tf, name)(tf.input([-1], tf.float32))]); raise SystemExit(name + ' traced: add it to the tests')\n"
        "    except RuntimeError as e:\n"
        "        assert 'not defined' in str(e), e\n"
        "rc = abi.lib().tfcuda_nvrtc_check(src.encode(), b''); assert rc == 0, abi.lib().tfcuda_last_error().decode()\n"
        "print('NVRTC-OK')\n") % (os.path.dirname(here), here)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert "NVRTC-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_unpinned_float_ops_match_numpy(tf_cuda):
    x, y = _inputs()
    outs = _program(tf_cuda)(x, y)
    names = ["rcp", "rsqrt", "frac", "modf", "rcp(negative)"]
    for name, o, w in zip(names, outs, _expected(x, y)):
        g = np.array(o.numpy)
        if name in ("frac", "modf"):  # results near 0 come from a cancellation: compare against the magnitude of the operand
            assert float(np.max(np.abs(g.astype(np.float64) - w.astype(np.float64)))) <= 4e-6, name
        else:
            _close(g, w, name)


@pytest.mark.gpu
def test_step_and_discard_in_generated_code(tfcuda_lib):
    """`step(edge, x)` and the `discard` keyword exist only at the code-generation level: a kernel in the emitter's shape, compiled and
    launched through the C-ABI (tfcuda_compile_kernels / tfcuda_launch)."""
    from tensorfrost_b200 import abi
    lib = tfcuda_lib
    n = 1000
    src = r'''
struct kernel_900001_args { uint* mem[2]; uint var[2]; };
extern "C" __global__ void __launch_bounds__(256) kernel_900001(const __grid_constant__ kernel_900001_args tf_a)
{
  uint* out_mem = tf_a.mem[0];
  TF_RO in_mem = tf_a.mem[1];
  int var_n = asint(tf_a.var[0]);
  uint var__kernel_block_offset = asuint(tf_a.var[1]);
  int block_id = (int)(blockIdx.x + var__kernel_block_offset);
  int index_0 = block_id * 256 + (int)threadIdx.x;
  if (index_0 < var_n) {
    float x = asfloat(in_mem[index_0]);
    if (x < -0.5f) { discard; }
    out_mem[index_0] = asuint(tf_step(0.25f, x) + 2.0f * tf_trunc(x * 3.5f));
  }
}
'''
    rec = abi.TFCudaKernelSource()
    rec.kernel_id, rec.entry, rec.source = 900001, b"kernel_900001", src.encode()
    rec.group[0], rec.group[1], rec.group[2] = 256, 1, 1
    rec.n_mem, rec.n_var, rec.library_op = 2, 2, 0
    abi.check(lib.tfcuda_compile_kernels(C.byref(rec), 1, b""), "compile")
    x = np.random.default_rng(5).uniform(-1, 1, n).astype(np.float32)
    din = abi.DeviceArray(x)
    dout = abi.DeviceArray(np.full(n, 7.0, np.float32))
    mem = (C.c_uint64 * 2)(dout.ptr, din.ptr)
    var = (C.c_uint32 * 2)(n, 0)
    abi.check(lib.tfcuda_launch(900001, mem, 2, var, 2, (n + 255) // 256), "launch")
    abi.check(lib.tfcuda_sync(), "sync")
    got = dout.get(np.float32)
    value = (x >= 0.25).astype(np.float32) + np.float32(2.0) * np.trunc(x * np.float32(3.5))
    want = np.where(x < -0.5, np.float32(7.0), value)  # discarded threads leave the output untouched
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_aliased_read_only_and_writable_binding_is_refused(tf_cuda):
    """Read-only bindings are `const uint* __restrict__` in emitted kernels; binding one device buffer both read-only and writable (the
    same TensorMemory passed as two program inputs, one of which the kernel writes) would make that promise false.  The dispatch refuses
    it with a message naming the switch (-DTF_NO_RESTRICT) instead of computing on possibly stale loads."""
    tf = tf_cuda

    def prog():
        a = tf.input([-1], tf.float32)
        b = tf.input(a.shape, tf.float32)
        i, = a.indices
        a[i] = b[i] * 2.0 + 1.0
        return a
    p = tf.compile(prog)
    x = tf.tensor(np.arange(64, dtype=np.float32))
    y = tf.tensor(np.ones(64, dtype=np.float32))
    out = p(x, y)
    assert np.array_equal(np.array(out.numpy), np.full(64, 3.0, np.float32))
    with pytest.raises(RuntimeError, match="TF_NO_RESTRICT"):
        p(x, x)
