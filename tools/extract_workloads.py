#!/usr/bin/env python3
"""Extract the reference's example PROGRAMS that BASELINE.json's configs name into build/workloads/ (git-ignored).

The configs are user programs shipped with the reference (examples/Simulation/fluid_simulation.ipynb cell 0,
examples/ML/NCA/nca.py).  They are the benchmark's INPUT, not code of this repository, so they are not stored in
git: this script copies them out of /root/reference at build time (build() runs it; the result travels to the GPU
box with the rest of build/).  Only mechanical edits are applied: the notebook cell loses its imports, its
`tf.initialize(tf.opengl)` line and its fixed resolution, which become parameters of the loader in
tensorfrost_b200/workloads.py.
"""
import json
import os
import re
import sys

REF = os.environ.get("TF_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "build", "workloads")


def main():
    if not os.path.isdir(os.path.join(REF, "examples")):
        print(f"[workloads] {REF} not present; using prebuilt {OUT} if any", file=sys.stderr)
        return
    os.makedirs(OUT, exist_ok=True)
    nb = json.load(open(os.path.join(REF, "examples", "Simulation", "fluid_simulation.ipynb")))
    cell0 = "".join(nb["cells"][0]["source"])
    lines = []
    for line in cell0.splitlines():
        if re.match(r"\s*(import |from )", line) or line.startswith("tf.initialize("):
            continue
        if re.match(r"M\s*=\s*\d+\s*$", line):
            line = "M = _M"
        if re.match(r"N\s*=\s*\d+\s*$", line):
            line = "N = _N"
        lines.append(line)
    text = "\n".join(lines) + "\n"
    assert "M = _M" in text and "N = _N" in text and "fluid = tf.compile(FluidTest)" in text
    with open(os.path.join(OUT, "fluid_program.py.txt"), "w") as f:
        f.write(text)
    nca = open(os.path.join(REF, "examples", "ML", "NCA", "nca.py")).read()
    with open(os.path.join(OUT, "nca_program.py.txt"), "w") as f:
        f.write(nca)
    # the two n-body programs of examples/Simulation/n-body-benchmark.py:16-65 (the file itself imports torch / jax and opens a window)
    src = open(os.path.join(REF, "examples", "Simulation", "n-body-benchmark.py")).read()
    start, end = src.index("def n_body():"), src.index("def n_body_torch(")
    body = src[start:end]
    assert "def n_body_loop():" in body and "tf.loop(N)" in body
    with open(os.path.join(OUT, "nbody_program.py.txt"), "w") as f:
        f.write(body)
    print(f"[workloads] extracted fluid + NCA + n-body programs into {OUT}")


if __name__ == "__main__":
    main()
