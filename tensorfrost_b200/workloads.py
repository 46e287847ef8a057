"""Loaders for the benchmark programs named by BASELINE.json's configs.

The programs are the reference's own example sources, extracted into build/workloads/ by tools/extract_workloads.py
(they are input data for the benchmark, not code of this repository).  Every loader takes the TensorFrost module to
trace with, so the SAME program runs on the oracle (tf.cpu) and on the CUDA backend.
Synthetic inputs follow SURVEY.md §8(d).
"""
import os
import types

import numpy as np

from . import REPO_ROOT

WORKLOAD_DIR = os.path.join(REPO_ROOT, "build", "workloads")


def _source(name):
    path = os.path.join(WORKLOAD_DIR, name)
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run tools/extract_workloads.py where the reference sources exist")
    return open(path).read(), path


# ---- C2: 2-D Eulerian fluid (examples/Simulation/fluid_simulation.ipynb, cell 0) -------------------------------------
def load_fluid(tf, n, m):
    """Compile the fluid step for an n x m grid; returns the compiled program
    fluid(vx, vy, pressure, density, mouse, params) -> [vx, vy, pressure, canvas, div, density, residual]."""
    text, path = _source("fluid_program.py.txt")
    ns = {"tf": tf, "np": np, "_N": int(n), "_M": int(m)}
    exec(compile(text, path, "exec"), ns)
    return ns["fluid"]


def fluid_inputs(n, m):
    """notebook cell 2 defaults with a constant 'mouse' source in the middle of the grid."""
    z = np.zeros((n, m), np.float32)
    mouse = np.array([m / 2, n / 2, 0.5, 0.5, 1.0], np.float32)
    params = np.array([1.0, 0.0, 1.0, 0.0, 0.999, 0.999], np.float32)
    return [z.copy(), z.copy(), z.copy(), z.copy(), mouse, params]


def fluid_step(fluid, state):
    """One step feeding outputs back: state = [vx, vy, pressure, density, mouse, params] (tensors or arrays)."""
    vx, vy, pressure, canvas, div, density, res = fluid(*state)
    return [vx, vy, pressure, density, state[4], state[5]], (canvas, div, res)


def fluid_parity_mouse(step, n, m):
    """A source that moves on a circle, so advection, the pressure solve and the canvas see non-trivial fields after a few steps."""
    a = 0.35 * step
    return np.array([m / 2 + 0.2 * m * np.cos(a), n / 2 + 0.2 * n * np.sin(a), 0.8 * np.cos(a + 1.0), 0.8 * np.sin(a + 1.0), 1.0], np.float32)


def fluid_parity_run(tf, n, m, steps, vorticity=0.0, program=None):
    """The parity scenario of the headline workload (tests/golden/make_golden_fluid.py on the oracle, tests/test_zz_fluid_gpu.py on the
    CUDA backend): `steps` steps from rest with the moving source, outputs fed back; returns [vx, vy, pressure, density, div, canvas] of the
    last step as numpy.  Notebook default parameters (vorticity confinement scale 0: with confinement on the program normalises the curl
    gradient, grad / (|grad| + 1e-5), which is discontinuous where the gradient vanishes and turns last-bit differences into O(1) ones)."""
    fluid = program if program is not None else load_fluid(tf, n, m)
    state = fluid_inputs(n, m)
    state[5] = np.array([1.0, vorticity, 1.0, 1.0, 0.999, 0.999], np.float32)
    div = canvas = None
    for s in range(steps):
        state[4] = fluid_parity_mouse(s, n, m)
        state, (canvas, div, _res) = fluid_step(fluid, state)
    return [np.array(t.numpy) for t in state[:4]] + [np.array(div.numpy), np.array(canvas.numpy)]


# ---- C5: neural cellular automata training (examples/ML/NCA/nca.py) ---------------------------------------------------
def load_nca(tf, batch_size, grid, pool_size=1024, train_steps=25, channel_n=12, quantize=True):
    """Exec the NCA example as a module with its size constants overridden; returns the module namespace
    (CAModel, CATrain, optimization_step, ...).  grid = TARGET_SIZE + 2*TARGET_PADDING.
    quantize=False replaces CAModel.quantize (round to 1/255 with a straight-through gradient, nca.py:58-59) by the identity: the
    smooth variant the parity tests use to pin gradients at 1e-3 (with the rounding on, an ulp upstream flips cells)."""
    text, path = _source("nca_program.py.txt")
    mod = types.ModuleType("nca_workload")
    mod.__file__ = path
    import sys
    sys.modules.setdefault("TensorFrost", tf)
    exec(compile(text, path, "exec"), mod.__dict__)
    mod.CHANNEL_N = channel_n
    mod.TARGET_PADDING = 8
    mod.TARGET_SIZE = grid - 2 * mod.TARGET_PADDING
    mod.BATCH_SIZE = batch_size
    mod.POOL_SIZE = pool_size
    mod.DEFAULT_TRAIN_STEPS = train_steps
    if not quantize:
        mod.CAModel.quantize = lambda self, Xstate: Xstate
    return mod


def nca_target(grid, seed=0):
    """Synthetic premultiplied RGBA target (there is no network for the emoji the example downloads)."""
    rng = np.random.default_rng(seed)
    img = rng.random((grid, grid, 4)).astype(np.float32)
    img[..., :3] *= img[..., 3:]
    return img


def nca_filters():
    sobel = np.array([[-1, -2, -1], [0, 0, 0], [1, 2, 1]], np.float32)
    laplace = np.array([[1, 2, 1], [2, -12, 2], [1, 2, 1]], np.float32)
    return np.stack([sobel, sobel.T, laplace], axis=0)


def nca_pool(pool_size, grid, channel_n):
    pool = np.zeros([pool_size, grid, grid, channel_n], np.float32)
    pool[:, grid // 2, grid // 2, 3:] = 1.0
    return pool


# ---- C3: the reference's n-body programs (extracted like the others); C4: one-line programs written here -----------------
def _nbody_namespace(tf):
    text, path = _source("nbody_program.py.txt")
    ns = {"tf": tf, "np": np}
    exec(compile(text, path, "exec"), ns)
    return ns


def compile_nbody(tf):
    """examples/Simulation/n-body-benchmark.py:16-34 (`n_body`: broadcast differences + tf.sum over the partner axis)."""
    return tf.compile(_nbody_namespace(tf)["n_body"])


def compile_nbody_loop(tf):
    """examples/Simulation/n-body-benchmark.py:36-65 (`n_body_loop`: one thread per body, explicit tf.loop over the partners)."""
    return tf.compile(_nbody_namespace(tf)["n_body_loop"])


def compile_matmul(tf):
    def prog():
        a = tf.input([-1, -1], tf.float32)
        b = tf.input([a.shape[1], -1], tf.float32)
        return a @ b
    return tf.compile(prog)


def compile_row_reductions(tf, n):
    def prog():
        a = tf.input([-1, n], tf.float32)
        return tf.sum(a), tf.max(a), tf.mean(a), tf.norm(a)
    return tf.compile(prog)


def compile_sort(tf, with_values=True):
    def prog():
        keys = tf.input([-1], tf.uint32)
        if with_values:
            values = tf.input([-1], tf.uint32)
            k, v = tf.sort.radix(keys, values)
            return k, v
        return tf.sort.radix(keys)
    return tf.compile(prog)
