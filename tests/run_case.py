"""Run named cases on ONE backend in a fresh process and dump their outputs.

The reference's backend is a process-global singleton (Backend/Backend.cpp:5-8), so the oracle (tf.cpu,
from oracle/_ref) and the CUDA backend never share a process: parity tests run this script once per backend
and compare the .npz files.

usage: python tests/run_case.py <cpu|cpu_fast|cuda> <out.npz> <case[:size[:seed]]> [...]
  cpu       oracle, compiled with "-O3 -fopenmp -include math.h" (no fast-math): the parity reference.
            `-include math.h` gives the generated C++ the float overload of abs() that the reference's author gets from
            MSVC's <cmath>; with g++ the generated `abs(x)` on a float otherwise binds to ::abs(int) and truncates.
  cpu_fast  oracle with the reference's default flags (-O3 -ffast-math -fopenmp): timing only
  cuda      the B200 backend
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


ORACLE_PARITY_FLAGS = "-O3 -fopenmp -include math.h"


def load_backend(which):
    if which in ("cpu", "cpu_fast"):
        ref = os.path.join(ROOT, "oracle", "_ref")
        if not os.path.isdir(os.path.join(ref, "TensorFrost")):
            raise SystemExit("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
        sys.path.insert(0, ref)
        import TensorFrost as tf
        tf.initialize(tf.cpu, ORACLE_PARITY_FLAGS if which == "cpu" else "")
        return tf
    if which == "cuda":
        import tensorfrost_b200
        return tensorfrost_b200.load(os.environ.get("TFCUDA_KERNEL_OPTIONS", ""))
    raise SystemExit(f"unknown backend {which}")


def main():
    which, out = sys.argv[1], sys.argv[2]
    specs = sys.argv[3:]
    tf = load_backend(which)
    import cases
    result = {}
    for spec in specs:
        parts = spec.split(":")
        name = parts[0]
        size = int(parts[1]) if len(parts) > 1 and parts[1] else None
        seed = int(parts[2]) if len(parts) > 2 else 0
        t0 = time.perf_counter()
        outs, _ = cases.run_case(tf, name, seed=seed, size=size)
        dt = time.perf_counter() - t0
        for k, o in enumerate(outs):
            result[f"{spec}/{k}"] = o
        result[f"{spec}/seconds"] = np.array(dt)
        print(f"[run_case] {which} {spec}: {len(outs)} outputs in {dt:.2f}s", flush=True)
    np.savez(out, **result)


if __name__ == "__main__":
    main()
