"""Thread coarsening on the device (overlay/Backend/CodeGen/Langs/CUDA.cpp): the fluid program emitted with 4 lanes per
thread (size threshold lowered so that the 256 x 256 program takes the lane code, edge paths included at 100 x 72) against the same
program emitted with one element per thread, in two processes.  The lane code is the same statements replicated per lane, so the
fields agree to rounding (a product shared by two lanes may be contracted into an FMA in one form and not in the other: the bar is
5e-5 of each field's maximum, half of the loosest parity bar against the reference); the CPU suite checks bit-identity on the host
(tests/test_coarsening_cpu.py)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
pytestmark = pytest.mark.gpu

SCRIPT = r'''
import sys, os, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import tensorfrost_b200
from tensorfrost_b200 import workloads
tf = tensorfrost_b200.load()
out = {}
keep = []  # programs must stay alive: the module's kernel registry holds raw pointers into them (Backend/KernelManager.cpp:4-8)
for n, m in ((256, 256), (100, 72)):
    fluid = workloads.load_fluid(tf, n, m)
    keep.append(fluid)
    state = [tf.cuda_tensor(a) for a in workloads.fluid_inputs(n, m)]
    for step in range(6):
        state[4] = tf.cuda_tensor(workloads.fluid_parity_mouse(step, n, m))
        state, (canvas, div, res) = workloads.fluid_step(fluid, state)
    for k, t in zip(("vx", "vy", "pressure", "density"), state[:4]):
        out[f"{n}x{m}_{k}"] = tf.cuda_numpy(t)
    out[f"{n}x{m}_canvas"] = tf.cuda_numpy(canvas)
    out[f"{n}x{m}_div"] = tf.cuda_numpy(div)
texts = [k[0][1] + k[0][2] for k in tf.get_all_generated_kernels()]
np.savez(sys.argv[1], **out)
print("LANES " + json.dumps([sum("lanes per thread" in t for t in texts), sum("tf_lane" in t for t in texts), len(texts)]))
''' % (ROOT, HERE)


def _run(tmp_path, tag, env_extra):
    script = tmp_path / "lanes.py"
    script.write_text(SCRIPT)
    out = tmp_path / f"{tag}.npz"
    env = dict(os.environ)
    env.update(env_extra)
    r = subprocess.run([sys.executable, str(script), str(out)], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    counts = json.loads([l for l in r.stdout.splitlines() if l.startswith("LANES ")][0][6:])
    return np.load(out), counts


def test_lane_code_matches_one_element_per_thread_on_the_device(tmp_path):
    plain, c0 = _run(tmp_path, "plain", {"TFCUDA_COARSEN": "0"})
    lanes, c1 = _run(tmp_path, "lanes", {"TFCUDA_COARSEN_MIN_ELEMENTS": "1"})
    assert c0[0] == 0, c0
    assert c1[0] >= 10 and c1[1] >= 1, c1  # most of the 2 x 15 kernels carry lanes; the 100 x 72 program has edge paths
    assert set(plain.files) == set(lanes.files)
    for k in plain.files:
        a, b = plain[k].astype(np.float64), lanes[k].astype(np.float64)
        assert a.shape == b.shape, k
        scale = max(np.abs(a).max(), 1e-30)
        assert np.isfinite(b).all() and np.abs(a - b).max() <= 5e-5 * scale, (k, float(np.abs(a - b).max()), float(scale))
