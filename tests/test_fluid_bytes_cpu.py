"""bench.py's closed form for the algorithmic bytes of a fluid step (`fluid_step_bytes`: the roofline numerator of the headline metric,
SURVEY.md §8d C2) recomputed from the program itself: the host program the reference generates for the 2048 x 2048 fluid step is traced in
codegen mode (nothing executes), every `tf.allocate` / input shape and every `tf.dispatch` is parsed, host loops are unrolled, and the
sizes of the tensors bound to each dispatch are summed, each once.  Must equal the closed form - and the count the runtime's profiler
reported live on the B200 (816,840,772 bytes, BENCH_r01.json / profiles/r02_bench.json)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = os.path.isdir(os.path.join(ROOT, "build", "tf_cuda", "TensorFrost")) and os.path.isdir(os.path.join(ROOT, "build", "workloads"))

DUMP = r'''
import sys
sys.path.insert(0, %r)
import tensorfrost_b200
from tensorfrost_b200 import workloads
tf = tensorfrost_b200.import_module()
tf.initialize(tf.codegen, "", tf.cuda_lang)
f = workloads.load_fluid(tf, int(sys.argv[1]), int(sys.argv[1]))
open(sys.argv[2], "w").write(f.get_main_function())
''' % ROOT


def dispatch_bytes(main_text, n):
    """Sum over dispatches (host `for` loops unrolled) of 4 bytes x elements of every bound tensor."""
    sizes = {"mouse": 5, "params": 6}
    for name in ("vx", "vy", "pressure", "density"):
        sizes[name] = n * n
    consts = {}
    total, launches = 0, 0
    stack = [1]  # loop multipliers
    for line in main_text.splitlines():
        line = line.strip()
        m = re.match(r"int (\w+) = (\d+);", line)
        if m:
            consts[m.group(1)] = int(m.group(2))
        m = re.match(r"TFTensor (\w+) = tf\.allocate\(\"[^\"]*\", \{([^}]*)\}", line)
        if m:
            dims = [d.replace("(uint)", "").strip() for d in m.group(2).split(",")]
            count = 1
            for d in dims:
                count *= consts[d] if d in consts else int(d)
            sizes[m.group(1)] = count
        m = re.match(r"for \(int \w+ = 0; \w+ < (\d+); \w+ \+= 1\)", line)
        if m:
            stack.append(stack[-1] * int(m.group(1)))
            stack.append(None)  # marks "opened by a for"
            continue
        if line == "{" and stack and stack[-1] is None:
            stack.pop()
            stack.append("open")
            continue
        if line == "}" and len(stack) > 1 and stack[-1] == "open":
            stack.pop()
            stack.pop()
            continue
        m = re.match(r"tf\.dispatch\((\d+), \{([^}]*)\},\s*\{([^}]*)\}", line)
        if m:
            mult = [s for s in stack if isinstance(s, int)][-1]
            names = [x.strip() for x in (m.group(2) + "," + m.group(3)).split(",") if x.strip()]
            total += mult * sum(4 * sizes[x] for x in names)
            launches += mult
    return total, launches


@pytest.mark.skipif(not HAVE, reason="CUDA-enabled module / extracted workloads not built here")
def test_closed_form_equals_the_dispatch_list(tmp_path):
    sys.path.insert(0, ROOT)
    import bench
    for n in (2048, 512):
        out = tmp_path / f"main_{n}.txt"
        r = subprocess.run([sys.executable, "-c", DUMP, str(n), str(out)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr[-2000:]
        total, launches = dispatch_bytes(out.read_text(), n)
        assert launches == 43, launches
        assert total == bench.fluid_step_bytes(n), (n, total, bench.fluid_step_bytes(n))
    assert bench.fluid_step_bytes(2048) == 816840772
