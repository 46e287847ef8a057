#!/usr/bin/env python3
"""Writes profiles/<name>: per-kernel SASS mnemonic counts of libtfcuda.so (the evidence that the matmul really is tcgen05 / TMEM / TMA,
the sort uses match.any, the n-body kernel packed f32x2 arithmetic, ...) plus an excerpt of the MMA loop.  Runs on the CPU box:
    python tools/sass_evidence.py profiles/r02_sass_library_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "MATCH.ANY", "REDG", "RED.E", "FFMA2", "FMUL2", "FADD2", "LDG.E.128", "STG.E.128",
        "ATOMS", "ATOMG", "REDUX", "MUFU.RSQ", "VOTE", "SHFL"]


def main():
    out_path = sys.argv[1]
    lib = os.path.join(ROOT, "tensorfrost_b200", "lib", "libtfcuda.so")
    txt = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    out = ["# SASS evidence for the hand-written kernels of libtfcuda.so",
           "# command: cuobjdump -sass tensorfrost_b200/lib/libtfcuda.so   (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -lineinfo -O3)",
           "# per kernel: instruction count and the counts of the mnemonics that identify the hardware path",
           "#   UTCHMMA = tcgen05.mma   LDTM = tcgen05.ld (TMEM -> registers)   UTMALDG = TMA tile load   UTCBAR = tcgen05.commit -> mbarrier",
           "#   SYNCS = mbarrier ops   MATCH.ANY = __match_any_sync   REDG / RED = red.global   FFMA2 / FMUL2 / FADD2 = packed f32x2 arithmetic",
           "#   LDG.E.128 / STG.E.128 = 128-bit global access   ATOMS / ATOMG = atomics   REDUX = __reduce_*_sync", ""]
    rows = []
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        ins = re.findall(r"/\*[0-9a-f]{4}\*/\s+([^;]+);", f)
        cnt = collections.Counter()
        for i in ins:
            parts = i.split()
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            for k in KEYS:
                if op.startswith(k):
                    cnt[k] += 1
        short = re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu_[0-9a-f]{8}", r"\1::", name)
        rows.append((short, len(ins), cnt))
    for name, n, cnt in sorted(rows):
        out.append(f"{name[:100]:100s} {n:6d} instr  " + "  ".join(f"{k}={v}" for k, v in sorted(cnt.items())))
    out += ["", "# excerpt: TMA loads, the MMA issue loop and the TMEM read-back of gemm_tf32_kernel<256, false> (tcgen05 kind::tf32, accumulators in TMEM)"]
    for f in funcs[1:]:
        if "gemm_tf32_kernelILi256ELb0" in f.split("\n", 1)[0]:
            lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
            idx = [i for i, l in enumerate(lines) if any(k in l for k in ("UTCHMMA", "LDTM", "UTMALDG", "UTCBAR"))]
            shown = set()
            for i in idx[:16]:
                for k in range(max(0, i - 1), min(len(lines), i + 2)):
                    if k not in shown:
                        shown.add(k)
                        out.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/", "", lines[k]).rstrip())
            break
    with open(out_path, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print("wrote", out_path, len(out), "lines")


if __name__ == "__main__":
    main()
