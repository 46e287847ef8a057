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
