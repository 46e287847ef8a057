"""CPU test of the CUDA emitter: every fused kernel of every parity case (and of the fluid benchmark program) is traced by
the CUDA-enabled module in codegen mode and must compile with NVRTC for sm_100a (tests/emit_all.py, own process because
the backend is a process-global singleton).  Needs the module built by tensorfrost_b200/build_overlay.sh."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_every_emitted_kernel_compiles_for_sm100a(tmp_path):
    import tensorfrost_b200
    try:
        tensorfrost_b200.module_path()
    except ImportError:
        pytest.skip("CUDA-enabled module not built here (build() needs the reference sources)")
    if not os.path.exists(os.path.join(ROOT, "build", "workloads", "fluid_program.py.txt")):
        pytest.skip("benchmark programs not extracted")
    r = subprocess.run([sys.executable, os.path.join(HERE, "emit_all.py")], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    tail = "\n".join(l for l in r.stdout.splitlines() if l.startswith("[emit]") or l.startswith("----"))
    assert r.returncode == 0, tail + r.stderr[-2000:]
    assert "0 cases failed" in r.stdout


def test_library_call_lowering_structure(tmp_path):
    """Which hand-written kernels a compiled program calls (codegen mode, no device): the weight gradients of a dense layer applied to
    an N-D field are ONE tfcuda_matmul_tn each (no [batch, K, N] intermediate and no batch reductions of it), tf.sort.radix is one
    library call whose scratch is not a program buffer, long row reductions are library calls, TFCUDA_LIBRARY=0 restores the generic loops."""
    import json
    import tensorfrost_b200
    try:
        tensorfrost_b200.module_path()
    except ImportError:
        pytest.skip("CUDA-enabled module not built here (build() needs the reference sources)")
    r = subprocess.run([sys.executable, os.path.join(HERE, "lowering_probe.py")], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    line = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
    assert r.returncode == 0 and line, r.stdout[-1500:] + r.stderr[-1500:]
    probe = json.loads(line[0][6:])
    dense = probe["autograd_batched_dense"]["library_calls"]
    # forward: 2 matmuls; backward: dX of both layers (2) + dW of both layers (2 x matmul_tn) + the user's x2.T @ t2 (matmul_tn)
    assert sum(c.startswith("matmul_tn") for c in dense) == 3, dense
    assert sum(c.startswith("matmul:") for c in dense) >= 3, dense
    sort_calls = probe["sort_pairs"]["library_calls"]
    assert len(sort_calls) == 1 and sort_calls[0].startswith("sort:1:32"), sort_calls
    assert probe["sort_pairs"]["main"].count("tf.allocate(\"lib_out") == 2  # sorted keys + sorted values, no scratch tensor
    assert [c.split(":")[0] for c in probe["matmul"]["library_calls"]] == ["matmul"]
    assert sum(c.startswith("reduce:") for c in probe["row_reductions"]["library_calls"]) == 4
    assert probe["row_reductions_generic"]["library_calls"] == []
