"""Launch recorder + graph replay (csrc/runtime.cu): between tfcuda_graph_begin / tfcuda_graph_end the dispatches
of a program are recorded and issued as one CUDA graph.  The argument bytes are baked into the graph nodes, so a replay must produce
BIT-IDENTICAL results to eager launches: checked here on the fluid program (43 dispatches, outputs fed back: the pool alternates
between two address sets -> exact hits after warm-up), on a program with host readbacks inside a loop (the recorder must flush at every
tf.read) and on an atomics program, against a process that runs the same programs with TFCUDA_GRAPH=0."""
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
import cases
tf = tensorfrost_b200.load()
out = {}
n = 256
fluid = workloads.load_fluid(tf, n, n)
state = [tf.cuda_tensor(a) for a in workloads.fluid_inputs(n, n)]
for step in range(8):
    state[4] = tf.cuda_tensor(workloads.fluid_parity_mouse(step, n, n))
    state, (canvas, div, res) = workloads.fluid_step(fluid, state)
for k, t in zip(("vx", "vy", "pressure", "density"), state[:4]):
    out["fluid_" + k] = tf.cuda_numpy(t)
out["fluid_canvas"] = tf.cuda_numpy(canvas)
for name in ("host_loop", "atomics", "sort_radix_u32"):
    os.environ["TFCUDA_LIBRARY"] = "0"
    prog = None
    for rep in range(3):   # the same program three times: the second and third executions can replay
        outs, prog = cases.run_case(tf, name, seed=3, program=prog)
    for k, o in enumerate(outs):
        out[f"{name}_{k}"] = o
np.savez(sys.argv[1], **out)
print("STATS " + json.dumps(tf.cuda_graph_stats()))
''' % (ROOT, HERE)


def _run(tmp_path, tag, graph):
    script = tmp_path / "replay.py"
    script.write_text(SCRIPT)
    out = tmp_path / f"{tag}.npz"
    env = dict(os.environ, TFCUDA_GRAPH="1" if graph else "0")
    r = subprocess.run([sys.executable, str(script), str(out)], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    import json
    stats = json.loads([l for l in r.stdout.splitlines() if l.startswith("STATS ")][0][6:])
    return np.load(out), stats


def test_replay_is_bit_identical_to_eager_and_hits_after_warmup(tmp_path):
    eager, s0 = _run(tmp_path, "eager", False)
    replay, s1 = _run(tmp_path, "replay", True)
    assert not s0["enabled"] and s0["replays"] == 0
    assert s1["enabled"] and s1["replays"] > 0, s1
    # 8 fluid steps: a chain is launched eagerly the first time its arguments are seen, becomes a graph the second time, and is
    # replayed unchanged from then on (outputs are fed back: two alternating address sets)
    assert s1["instantiated"] >= 1 and s1["replays"] >= s1["instantiated"], s1
    assert set(eager.files) == set(replay.files)
    for k in eager.files:
        a, b = eager[k], replay[k]
        assert a.shape == b.shape and a.dtype == b.dtype, k
        if "atomics" in k and a.dtype.kind == "f":
            # float atomics are order-nondeterministic in EITHER mode (same bar as the parity case)
            np.testing.assert_allclose(a, b, rtol=1e-4, atol=1e-5, err_msg=k)
        else:
            assert np.array_equal(a.view(np.uint32) if a.dtype.itemsize == 4 else a, b.view(np.uint32) if b.dtype.itemsize == 4 else b), f"{k}: replay differs from eager"
