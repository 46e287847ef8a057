"""CPU test: the CUDA C++ the emitter produces (the exact text NVRTC compiles for sm_100a) is compiled by g++ through a host shim and
EXECUTED on the host, for every parity case, for the 10-step fluid scenario and for the NCA grad program (kernels that use tf.group_barrier get one host thread
per CUDA thread of a block and a std::barrier), and the results are compared with the golden outputs of the reference's own C++/OpenMP backend (tests/cpu_sim/run_sim.py; test infrastructure only).

This pins, without a GPU, everything about an emitted program that is not hardware: the kernel wrapper and argument block, binding
and variable order, block / thread index arithmetic, tail guards, the prelude's helper semantics (min/max/sign/pcg/... quirks included),
atomics, host loops and readbacks.  On this container every case comes out bit-identical to the reference; the assertion below is the
case's own tolerance (the host compiler may differ between boxes).  The GPU suite then only adds what the hardware adds."""
import os
import subprocess
import sys

import numpy as np
import pytest

import cases

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
FLUID_NAMES = ["vx", "vy", "pressure", "density", "div", "canvas"]
# cases whose programs change under the library lowering (tests/test_parity_gpu.py USES_LIBRARY + the weight-gradient case)
LIBRARY_CASES = ["row_reductions", "int_reductions", "prefix_sum", "sort_radix_u32", "sort_radix_f32", "sort_radix_i32", "matmul", "qr_inverse",
                 "autograd_mlp", "autograd_batched_dense"]


def _golden(name):
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    outs = []
    while f"out{len(outs)}" in g:
        outs.append(g[f"out{len(outs)}"])
    return int(g["size"]), int(g["seed"]), outs


def test_emitted_kernels_reproduce_the_reference_on_the_host(tmp_path):
    import tensorfrost_b200
    try:
        tensorfrost_b200.module_path()
    except ImportError:
        pytest.skip("CUDA-enabled module not built here (build() needs the reference sources)")
    if not os.path.exists(os.path.join(ROOT, "build", "workloads", "fluid_program.py.txt")):
        pytest.skip("benchmark programs not extracted")
    names = sorted(cases.CASES)
    specs = {}
    for n in names:
        size, seed, _ = _golden(n)
        specs[n] = f"{n}:{size}:{seed}"
    fluid = np.load(os.path.join(GOLDEN, "fluid_128.npz"))
    fluid_spec = f"fluid:{int(fluid['n'])}:{int(fluid['n'])}:{int(fluid['steps'])}"
    out = str(tmp_path / "sim.npz")
    library_specs = {n: specs[n] + ":library" for n in LIBRARY_CASES}
    # four simulator processes side by side (each traces, builds with g++ and runs its share of the programs)
    generic = list(specs.values())
    groups = [generic[0::2], generic[1::2], [fluid_spec, "nca", "nca:library"], list(library_specs.values())]
    procs = []
    for i, group in enumerate(groups):
        out_i = str(tmp_path / f"sim{i}.npz")
        procs.append((out_i, subprocess.Popen([sys.executable, os.path.join(HERE, "cpu_sim", "run_sim.py"), out_i] + group, cwd=str(tmp_path),
                                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    got = {}
    for out_i, proc in procs:
        _, err = proc.communicate(timeout=1200)
        assert proc.returncode == 0, err[-3000:]
        with np.load(out_i) as z:
            got.update({k: z[k] for k in z.files})
    exact = 0
    for n in names:
        _, _, want = _golden(n)
        have = [got[f"{specs[n]}/{k}"] for k in range(len(want))]
        cases.compare(cases.CASES[n], have, want)
        exact += all(np.array_equal(np.ascontiguousarray(a).view(np.uint8), np.ascontiguousarray(b).view(np.uint8)) for a, b in zip(have, want))
    # the same programs traced with the LIBRARY lowering: marker kernels, role -> binding maps, axis conventions, the sort call without a
    # scratch tensor, the matmul VJP rewritten as one contraction (tfcuda_matmul_tn) - with the numpy restatement standing in for the
    # hand-written kernels, so this pins the lowering's structure and arithmetic, not the kernels (those are GPU tests)
    for n in LIBRARY_CASES:
        _, _, want = _golden(n)
        cases.compare(cases.CASES[n], [got[f"{library_specs[n]}/{k}"] for k in range(len(want))], want)
    for k, name in enumerate(FLUID_NAMES):
        a, b = got[f"{fluid_spec}/{k}"].astype(np.float64), fluid[name].astype(np.float64)
        assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max(), f"fluid {name}"
    # the two programs of the data-parallel NCA step (grad: 83 emitted kernels - 3 CA steps, autodiff, float atomics, and the reference's
    # out-of-range neighbour read, which the zero guard bands make deterministic; apply: 11 kernels) against the reference's gradients,
    # loss, state and 3-iteration loss sequence
    nca = np.load(os.path.join(GOLDEN, "nca_step.npz"))
    scale = np.abs(nca["flat0"][:-1]).max()
    for spec in ("nca", "nca:library"):  # generic lowering; library lowering (matmul / matmul_tn / reduce calls behind the numpy stand-ins)
        flat, state = got[f"{spec}/2"], got[f"{spec}/3"]
        assert flat.shape == nca["flat0"].shape
        assert np.abs(flat[:-1].astype(np.float64) - nca["flat0"][:-1]).max() <= 1e-5 * scale, spec
        assert abs(float(flat[-1]) - float(nca["split_losses"][0])) <= 1e-6, spec
        assert np.abs(state - nca["state0"]).max() <= 2.0 / 255.0 + 1e-6 and np.mean(np.abs(state - nca["state0"]) > 1e-6) < 0.02, spec
        # three whole training iterations (grad program -> apply program with norm clipping + Adam -> next grad program): the loss sequence
        np.testing.assert_allclose(got[f"{spec}/losses"], nca["split_losses"], rtol=1e-5, err_msg=spec)
    flat = got["nca/2"]
    print(f"{exact} of {len(names)} cases bit-identical to the reference; NCA gradients max |diff| "
          f"{np.abs(flat[:-1].astype(np.float64) - nca['flat0'][:-1]).max():.1e}")


ODD_SPECS = ["wave:37:3", "wave:1:4", "math_ops:1:5", "math_ops:257:6", "int_ops:33:7", "prefix_sum:1000:8", "atomics:777:9", "reshape_reduce:3:10",
             "control_flow:130:11", "split_merge:8:12", "host_loop:17:13", "nbody:65:14", "matmul:33:15", "int_reductions:7:16"]


def test_emitted_kernels_match_a_live_reference_run_at_odd_sizes(tmp_path):
    """Sizes that are not multiples of the block shape (tail guards, one-element tensors, partial last blocks) and other seeds than the
    fixtures: the reference itself (oracle/_ref, strict flags) and the host execution of the emitted CUDA text, same seeded inputs."""
    import tensorfrost_b200
    try:
        tensorfrost_b200.module_path()
    except ImportError:
        pytest.skip("CUDA-enabled module not built here")
    if not os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "TensorFrost")):
        pytest.skip("oracle/_ref (the reference module) is not built here")
    ref, sim = str(tmp_path / "ref.npz"), str(tmp_path / "sim.npz")
    for cmd in ([sys.executable, os.path.join(HERE, "run_case.py"), "cpu", ref] + ODD_SPECS,
                [sys.executable, os.path.join(HERE, "cpu_sim", "run_sim.py"), sim] + ODD_SPECS):
        r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-3000:]
    want_all, got_all = np.load(ref), np.load(sim)
    for spec in ODD_SPECS:
        want = []
        while f"{spec}/{len(want)}" in want_all:
            want.append(want_all[f"{spec}/{len(want)}"])
        assert want, spec
        cases.compare(cases.CASES[spec.split(":")[0]], [got_all[f"{spec}/{k}"] for k in range(len(want))], want)
