"""Generate the golden fixtures in tests/golden/ by running every case of tests/cases.py on the REFERENCE's own
C++/OpenMP backend (the unmodified module built by oracle/build_ref.sh, compiled without fast-math).

The reference holds no golden vectors of its own, so these outputs — produced by the reference
itself in the build container, where /root/reference exists — are what pins parity on the GPU box, where it
does not.  Fixtures are keyed by (case, size, seed); inputs are NOT stored, they are regenerated from the seed.

usage: python tests/golden/make_golden.py [case ...]
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
sys.path.insert(0, TESTS)
import cases  # noqa: E402

# (size, seed) per case for the committed fixture: small enough to keep the repository light
GOLDEN = {
    "wave": (96, 0), "math_ops": (1024, 0), "int_ops": (1024, 0), "pcgf_random": (2048, 0), "control_flow": (1024, 0),
    "reshape_reduce": (10, 0), "row_reductions": (2048, 0), "int_reductions": (150, 0), "prefix_sum": (3000, 0),
    "split_merge": (128, 0), "sort_radix_u32": (1 << 14, 0), "sort_radix_f32": (1 << 14, 0), "sort_radix_i32": (1 << 14, 0),
    "sort_bitonic_u32": (1 << 12, 0), "atomics": (20000, 0), "scatter_matmul": (48, 0), "matmul": (160, 0), "qr_inverse": (5, 0),
    "nbody": (512, 0), "nbody_loop": (512, 0), "host_loop": (200, 0), "autograd_mlp": (32, 0),
    "autograd_batched_dense": (6, 0),
}


def main():
    names = sys.argv[1:] or list(GOLDEN)
    for name in names:
        size, seed = GOLDEN[name]
        spec = f"{name}:{size}:{seed}"
        tmp = os.path.join("/tmp", f"golden_{name}.npz")
        subprocess.run([sys.executable, os.path.join(TESTS, "run_case.py"), "cpu", tmp, spec], check=True, cwd="/tmp",
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        data = np.load(tmp)
        outs = {}
        k = 0
        while f"{spec}/{k}" in data:
            outs[f"out{k}"] = data[f"{spec}/{k}"]
            k += 1
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), size=np.array(size), seed=np.array(seed), **outs)
        print(f"{name}: {k} outputs, {os.path.getsize(os.path.join(HERE, name + '.npz')) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
