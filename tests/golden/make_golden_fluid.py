"""Golden fixture for the headline workload: the 2-D fluid program (examples/Simulation/fluid_simulation.ipynb, extracted by
tools/extract_workloads.py) run for STEPS steps on the REFERENCE's C++/OpenMP backend (oracle/_ref, strict flags: no fast-math),
feeding outputs back (scenario: tensorfrost_b200.workloads.fluid_parity_run).  Stored: vx, vy, pressure, density,
div and canvas after the last step.

The reference has no test of this program, so this run is the pin.  The script also prints how far the same run moves when the
oracle's host compiler may contract a*b+c into FMAs (what nvcc does on the GPU) and with the reference's default -ffast-math: that
spread calibrates the tolerance of tests/test_zz_fluid_gpu.py (measured here: <= 1.4e-6 of max|field| for vx/vy/pressure/density,
<= 8.5e-6 for div and 2e-6 for canvas under either variant).

usage: python tests/golden/make_golden_fluid.py            (writes tests/golden/fluid_<N>.npz)
       python tests/golden/make_golden_fluid.py run <flags-name> <out.npz> [n m steps]     (one backend configuration per process)
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
N = 128
STEPS = 10
FLAGS = {
    "strict": "-O3 -fopenmp -include math.h",
    "fma": "-O3 -fopenmp -include math.h -mfma -ffp-contract=fast",
    "fast": "",  # the reference's default: -O3 -ffast-math -fopenmp
}
NAMES = ["vx", "vy", "pressure", "density", "div", "canvas"]


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "run":
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        sys.path.insert(0, ROOT)
        import TensorFrost as tf
        tf.initialize(tf.cpu, FLAGS[sys.argv[2]])
        from tensorfrost_b200 import workloads
        n, m, steps = (int(v) for v in sys.argv[4:7]) if len(sys.argv) >= 7 else (N, N, STEPS)
        outs = workloads.fluid_parity_run(tf, n, m, steps)
        np.savez(sys.argv[3], **dict(zip(NAMES, outs)))
        return
    results = {}
    for name in FLAGS:
        tmp = f"/tmp/golden_fluid_{name}.npz"
        subprocess.run([sys.executable, os.path.abspath(__file__), "run", name, tmp], check=True, cwd="/tmp", stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        results[name] = np.load(tmp)
    strict = results["strict"]
    for other in ("fma", "fast"):
        for k in NAMES:
            a, b = strict[k].astype(np.float64), results[other][k].astype(np.float64)
            scale = max(np.abs(a).max(), 1e-30)
            print(f"{other:6s} vs strict  {k:9s} max|diff|/max|ref| = {np.abs(a - b).max() / scale:.3e}   (max|ref| = {scale:.3e})")
    np.savez_compressed(os.path.join(HERE, f"fluid_{N}.npz"), n=np.array(N), steps=np.array(STEPS), **{k: strict[k] for k in NAMES})
    print("wrote", os.path.join(HERE, f"fluid_{N}.npz"), os.path.getsize(os.path.join(HERE, f"fluid_{N}.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
