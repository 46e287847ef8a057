"""CPU test of the emitter's thread coarsening (overlay/Backend/CodeGen/Langs/CUDA.cpp): large constant-shape straight-line kernels are
emitted with several "lanes" per CUDA thread (the IR's block is `factor` times the launched one along one dimension, the body is
replicated statement by statement).  The emitted text is executed on the host (tests/cpu_sim) with the size threshold lowered so that
small programs take the lane code, and compared with the same programs emitted WITHOUT coarsening - which the other host-execution
tests pin to the reference bit for bit.  Shapes that are not multiples of the enlarged block exercise the lane-by-lane path of blocks
that cross the edge of the dispatch."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
# fluid n x m x steps: 100 x 72 and 33 x 250 cross the edge of the enlarged block (32 x 32) in one / both dimensions, 64 x 96 divides it
SPECS = ["fluid:100:72:3", "fluid:33:250:2", "fluid:64:96:3"]


def _sim(tmp_path, tag, env_extra, specs):
    env = dict(os.environ)
    env.update(env_extra)
    out = str(tmp_path / f"{tag}.npz")
    return out, subprocess.Popen([sys.executable, os.path.join(HERE, "cpu_sim", "run_sim.py"), out] + specs, cwd=str(tmp_path), env=env,
                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)


def _ready():
    import tensorfrost_b200
    try:
        tensorfrost_b200.module_path()
    except ImportError:
        pytest.skip("CUDA-enabled module not built here (build() needs the reference sources)")
    if not os.path.exists(os.path.join(ROOT, "build", "workloads", "fluid_program.py.txt")):
        pytest.skip("benchmark programs not extracted")


def test_lane_code_is_bit_identical_to_one_element_per_thread(tmp_path):
    _ready()
    runs = {"off": _sim(tmp_path, "off", {"TFCUDA_COARSEN": "0"}, SPECS),
            "x4": _sim(tmp_path, "x4", {"TFCUDA_COARSEN_MIN_ELEMENTS": "1"}, SPECS),
            "x2": _sim(tmp_path, "x2", {"TFCUDA_COARSEN_MIN_ELEMENTS": "1", "TFCUDA_COARSEN": "2"}, SPECS),
            # the fallback for bodies the replication does not accept: the lanes one after the other around the untouched text
            "loop": _sim(tmp_path, "loop", {"TFCUDA_COARSEN_MIN_ELEMENTS": "1", "TFCUDA_COARSEN_FORCE_LOOP": "1"}, SPECS)}
    got = {}
    for tag, (out, proc) in runs.items():
        _, err = proc.communicate(timeout=1200)
        assert proc.returncode == 0, err[-3000:]
        with np.load(out) as z:
            got[tag] = {k: z[k] for k in z.files}
    for spec in SPECS:
        assert got["off"][f"{spec}/coarsened"][0] == 0
        assert got["loop"][f"{spec}/lane_loops"] >= 5 and got["loop"][f"{spec}/coarsened"][0] == 0, spec
        for tag in ("x4", "x2", "loop"):
            lanes, with_edge_path, kernels = got[tag][f"{spec}/coarsened"]
            if tag != "loop":
                assert lanes >= 5, (spec, tag, lanes, kernels)  # the full-resolution stencils at least (coarser multigrid levels may be too short)
                assert (with_edge_path > 0) == (spec != "fluid:64:96:3"), (spec, tag, with_edge_path)
            for k in range(6):  # vx, vy, pressure, density, div, canvas
                a, b = got[tag][f"{spec}/{k}"], got["off"][f"{spec}/{k}"]
                assert a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)), (spec, tag, k)


def test_which_kernels_of_the_benchmark_programs_get_lanes(tmp_path):
    """At the benchmark size (2048 x 2048) the ten stencil kernels of the 2048 and 1024 grids are coarsened; the kernels of the 512
    grid (too few blocks left to fill the machine) and the two long kernels (advection: 84 data-dependent gathers; final projection:
    no gain measured) are not; the
    kernel text states the block it must be launched with, which differs from the block of the host program's dispatch."""
    _ready()
    script = r"""
import re, sys
sys.path.insert(0, %r)
import tensorfrost_b200
from tensorfrost_b200 import workloads
tf = tensorfrost_b200.import_module()
tf.initialize(tf.codegen, "", tf.cuda_lang)
fluid = workloads.load_fluid(tf, 2048, 2048)
host = fluid.compiled_code()
for k in tf.get_all_generated_kernels():
    text = k[0][1] + k[0][2]
    kid = int(re.search(r"void (?:__launch_bounds__\(\d+\) )?kernel_(\d+)\(", text).group(1))
    block = re.search(r"// tfcuda_block: (\d+) (\d+) (\d+)", text).groups()
    ir_block = re.search(r"tf\.dispatch\(%%d,.*\{([^{}]*)\}\);" %% kid, host).group(1).replace(" ", "")
    print("KERNEL", kid, int("lanes per thread" in text), ",".join(block), ir_block, len(text.splitlines()))
from tensorfrost_b200 import abi
unit = "\n".join(k[0][1] + k[0][2] for k in tf.get_all_generated_kernels())
rc = abi.lib().tfcuda_nvrtc_check(unit.encode(), b"")
print("NVRTC", rc, abi.lib().tfcuda_last_error().decode(errors="replace")[:2000] if rc else "")
""" % ROOT
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-3000:]
    assert "NVRTC 0" in r.stdout, r.stdout[-3000:]  # the lane code compiles for sm_100a
    rows = [l.split() for l in r.stdout.splitlines() if l.startswith("KERNEL")]
    assert len(rows) == 15
    lanes = [row for row in rows if row[2] == "1"]
    assert len(lanes) == 10, rows
    for _, kid, _, block, ir_block, _ in lanes:
        assert block == "32,8,1" and ir_block == "32,32", (kid, block, ir_block)
    for _, kid, flag, block, ir_block, n_lines in rows:
        if flag == "0":
            assert (block + ",1,1").split(",")[:len(ir_block.split(","))] == ir_block.split(","), (kid, block, ir_block)


def test_nca_training_step_with_lanes_reproduces_the_reference(tmp_path):
    """The NCA grad program (float atomics, random masks, 9-tap filter loops, 3 CA steps + autodiff) with 62 of its 83 kernels carrying
    lanes - 21 of them with the lane-by-lane edge path, the filter kernels with their tap loop shared by the lanes - against the
    reference's own gradients, loss and 3-iteration loss sequence (tests/golden/nca_step.npz)."""
    _ready()
    out, proc = _sim(tmp_path, "nca", {"TFCUDA_COARSEN_MIN_ELEMENTS": "1"}, ["nca"])
    _, err = proc.communicate(timeout=1200)
    assert proc.returncode == 0, err[-3000:]
    got = np.load(out)
    lanes, with_edge_path, kernels = got["nca/coarsened"]
    assert lanes >= 40 and with_edge_path >= 1, (lanes, with_edge_path, kernels)
    nca = np.load(os.path.join(HERE, "golden", "nca_step.npz"))
    flat = got["nca/2"]
    scale = np.abs(nca["flat0"][:-1]).max()
    assert np.abs(flat[:-1].astype(np.float64) - nca["flat0"][:-1]).max() <= 1e-5 * scale
    assert abs(float(flat[-1]) - float(nca["split_losses"][0])) <= 1e-6
    np.testing.assert_allclose(got["nca/losses"], nca["split_losses"], rtol=1e-5)
