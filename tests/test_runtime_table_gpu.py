"""The drop-in boundary, executed: programs traced in CODEGEN mode, their reference-generated host
code compiled with g++ and called with the `TFRuntime` callback table that libtfcuda.so exports (`tfcuda_runtime()`): alloc, dealloc,
readback (tf.read in host loops), writeback, dispatch and region all run through the C-ABI, with no TensorFrost backend in the loop
(tests/standalone/run_standalone.py).  Results are compared with the reference's golden fixtures under each case's bar."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
RUNNER = os.path.join(HERE, "standalone", "run_standalone.py")
# elementwise, integer, control flow, host loop with tf.read, every atomic family, reshape views, the generic 13-kernel radix sort
TABLE_CASES = ["wave", "int_ops", "control_flow", "host_loop", "atomics", "split_merge", "sort_radix_u32", "prefix_sum"]


def _specs():
    specs = {}
    for name in TABLE_CASES:
        g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        specs[name] = f"{name}:{int(g['size'])}:{int(g['seed'])}"
    return specs


def test_standalone_programs_build_without_a_device(tmp_path):
    """CPU: trace in codegen mode, emit, g++ the host program against the ABI structs - everything up to the first device call."""
    out = tmp_path / "dry.npz"
    r = subprocess.run([sys.executable, RUNNER, "--dry", str(out), "wave", "host_loop"], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    d = np.load(out)
    assert int(d["wave/kernels"]) == 1 and int(d["host_loop/kernels"]) >= 2


@pytest.fixture(scope="module")
def standalone_outputs(tmp_path_factory):
    out = tmp_path_factory.mktemp("standalone") / "out.npz"
    specs = _specs()
    r = subprocess.run([sys.executable, RUNNER, str(out)] + list(specs.values()), cwd=str(out.parent), capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return np.load(out), specs


@pytest.mark.gpu
@pytest.mark.parametrize("name", TABLE_CASES)
def test_program_through_tfcuda_runtime_matches_golden(standalone_outputs, name):
    outs, specs = standalone_outputs
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    want, got = [], []
    while f"out{len(want)}" in g:
        want.append(g[f"out{len(want)}"])
    while f"{specs[name]}/{len(got)}" in outs:
        got.append(outs[f"{specs[name]}/{len(got)}"])
    cases.compare(cases.CASES[name], got, want)
