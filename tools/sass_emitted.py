"""Static SASS statistics of EMITTED kernels, without a GPU: traces the fluid benchmark program (2048 x 2048) and the per-rank NCA
grad program (batch 32 of 128 x 128, 25 steps) in codegen mode under three emitter settings, compiles every kernel for sm_100a with
nvcc and counts instructions with cuobjdump.  Used for the round-2 emitter work that could not be timed on hardware any more
(profiles/r02_sass_emitter_lanes.txt): instructions per ELEMENT, registers, and whether the loads of a thread are issued as one batch.

usage: python tools/sass_emitted.py [fluid] [nca] > profiles/<file>.txt         (each setting is traced in its own subprocess)"""
import collections
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETTINGS = [("round-2 measured build (no range fact, 1 element per thread)", {"TFCUDA_ASSUME": "0", "TFCUDA_COARSEN": "0"}),
            ("+ TF_ASSUME(block_id >= 0)", {"TFCUDA_COARSEN": "0"}),
            ("+ 4 lanes per thread (default)", {})]

TRACE = r"""
import json, os, sys
sys.path.insert(0, %r)
import tensorfrost_b200
from tensorfrost_b200 import workloads, nca_dp
tf = tensorfrost_b200.import_module()
tf.initialize(tf.codegen, "", tf.cuda_lang)
which = sys.argv[1]
if which == "fluid":
    keep = workloads.load_fluid(tf, 2048, 2048)
else:
    nca = workloads.load_nca(tf, 32, 128, pool_size=128)
    g, a, m, shapes = nca_dp.build_programs(tf, nca, 25)
    keep = tf.compile(g)
json.dump([k[0][1] + k[0][2] for k in tf.get_all_generated_kernels()], open(sys.argv[2], "w"))
""" % ROOT


def sass_stats(text, work):
    src = os.path.join(work, "k.cu")
    with open(src, "w") as f:
        f.write(open(os.path.join(ROOT, "tensorfrost_b200", "csrc", "prelude.cuh")).read() + "\n" + text)
    r = subprocess.run(["nvcc", "-arch=sm_100a", "-cubin", "-o", os.path.join(work, "k.cubin"), src, "-Xptxas", "-v"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-2000:])
    regs = int(re.search(r"Used (\d+) registers", r.stderr).group(1))
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(work, "k.cubin")], capture_output=True, text=True).stdout
    ops = [m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass, re.M)]
    ops = [o.split(".")[0] for o in ops if not o.startswith("NOP")]
    mem = [o for o in ops if o in ("LDG", "STG", "REDG", "ATOMG")]
    runs = []
    for o in mem:
        if runs and runs[-1][0] == o:
            runs[-1][1] += 1
        else:
            runs.append([o, 1])
    return len(ops), regs, collections.Counter(ops), " ".join(f"{n}x{o}" for o, n in runs)


def main():
    which = [a for a in sys.argv[1:] if a in ("fluid", "nca")] or ["fluid", "nca"]
    work = tempfile.mkdtemp(prefix="sass_emitted_")
    for w in which:
        per_setting = []
        for label, env_extra in SETTINGS:
            env = dict(os.environ)
            env.update(env_extra)
            out = os.path.join(work, f"{w}.json")
            r = subprocess.run([sys.executable, "-c", TRACE, w, out], env=env, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(r.stderr[-3000:])
            per_setting.append(json.load(open(out)))
        n = len(per_setting[0])
        print(f"== {w}: {n} kernels; columns per setting: SASS instructions per element (instructions / lanes), registers, memory instruction order")
        for label, _ in SETTINGS:
            print(f"   setting: {label}")
        # NCA repeats the same 25 CA steps: report each distinct (line count, lanes) class once, with its multiplicity
        classes = collections.OrderedDict()
        for i in range(n):
            texts = [s[i] for s in per_setting]
            if "__global__" not in texts[0]:
                continue
            key = tuple(len(t.splitlines()) for t in texts) if w == "nca" else i
            classes.setdefault(key, []).append(i)
        for key, members in classes.items():
            i = members[0]
            cells = []
            for s in per_setting:
                text = s[i]
                lanes = int(re.search(r"// (\d+) lanes per thread", text).group(1)) if "lanes per thread" in text else 1
                total, regs, mix, order = sass_stats(text, work)
                cells.append(f"{total / lanes:6.1f} instr/elem ({total} / {lanes}), {regs} regs, {order}")
            name = re.search(r"void (?:__launch_bounds__\([\d, ]+\) )?(kernel_\d+)\(", per_setting[0][i]).group(1)
            print(f"{name}" + (f" (x{len(members)} kernels of this form)" if len(members) > 1 else ""))
            for c in cells:
                print("    " + c)


if __name__ == "__main__":
    main()
