"""GPU parity of the HEADLINE workload (BASELINE.json configs[1]): the 2-D fluid program on the CUDA backend vs the reference's own
C++/OpenMP backend, 10 steps from rest with a moving source, outputs fed back (tensorfrost_b200.workloads.fluid_parity_run).

Two pins, as in tests/test_parity_gpu.py: the committed fixture tests/golden/fluid_128.npz (made by tests/golden/make_golden_fluid.py
on the oracle with strict flags) and, when oracle/_ref travelled to the box, a live oracle run at a rectangular size.

Bars (relative to each field's max magnitude after the 10 steps): 2e-5 for vx / vy / pressure / density / canvas, 1e-4 for div (a
difference of neighbouring velocities, its own scale is 10x smaller).  Calibration: the oracle itself moves by <= 2.2e-6 (fields),
<= 8.5e-6 (div), <= 5.2e-6 (canvas) between strict flags, FMA contraction (-mfma -ffp-contract=fast: what nvcc does) and the reference's
default -ffast-math (numbers printed by make_golden_fluid.py); the bars leave ~10x on top of that for the GPU's own expf / sqrtf / division
rounding over 10 steps of a multigrid solve.  north_star's single-operation bar is 1e-5.  The emitted kernels themselves reproduce this fixture
bit for bit when executed on the host (tests/test_emitted_cuda_on_host.py), so whatever this test sees is the hardware's arithmetic.
(File name: collected last, so the rest of the GPU suite is reported before this long-running scenario.)"""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HAVE_ORACLE = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "TensorFrost"))
NAMES = ["vx", "vy", "pressure", "density", "div", "canvas"]
TOL = {"vx": 2e-5, "vy": 2e-5, "pressure": 2e-5, "density": 2e-5, "div": 1e-4, "canvas": 2e-5}

pytestmark = pytest.mark.gpu


def _compare(got, want):
    for name, g in zip(NAMES, got):
        w = want[name]
        assert g.shape == w.shape, f"{name}: shape {g.shape} vs {w.shape}"
        assert np.isfinite(g).all(), f"{name}: non-finite values"
        scale = max(float(np.abs(w).max()), 1e-30)
        err = float(np.abs(g.astype(np.float64) - w.astype(np.float64)).max()) / scale
        assert err <= TOL[name], f"fluid {name}: max error {err:.3e} of max|ref| exceeds {TOL[name]:.0e}"
        assert float(np.abs(w).max()) > 1e-3, f"{name}: the reference field is trivially zero, the scenario pins nothing"


def test_fluid_10_steps_match_golden(tf_cuda):
    from tensorfrost_b200 import workloads
    g = np.load(os.path.join(HERE, "golden", "fluid_128.npz"))
    n, steps = int(g["n"]), int(g["steps"])
    _compare(workloads.fluid_parity_run(tf_cuda, n, n, steps), g)


@pytest.mark.skipif(not HAVE_ORACLE, reason="oracle/_ref (the reference module) is not present")
def test_fluid_matches_live_oracle_rectangular(tf_cuda, tmp_path):
    from tensorfrost_b200 import workloads
    n, m, steps = 96, 160, 6
    out = str(tmp_path / "fluid_oracle.npz")
    r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "make_golden_fluid.py"), "run", "strict", out, str(n), str(m), str(steps)],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    _compare(workloads.fluid_parity_run(tf_cuda, n, m, steps), np.load(out))
