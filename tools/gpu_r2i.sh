#!/usr/bin/env bash
# Round 2, short 1-GPU call: synccheck / racecheck of the generic (emitted) radix sort, register-cap experiments on the fluid step,
# the reference's n-body loop program with approximate division / square root.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2i; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
cat > "$OUT/generic_sort.py" <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
os.environ["TFCUDA_LIBRARY"] = "0"; os.environ["TFCUDA_GRAPH"] = os.environ.get("TFCUDA_GRAPH", "0")
import numpy as np
import tensorfrost_b200
tf = tensorfrost_b200.load()
import cases
outs, _ = cases.run_case(tf, "sort_radix_u32", seed=5, size=20003)   # not a multiple of any block size: tail blocks are partially outside
k, v = outs
assert np.all(k[1:] >= k[:-1])
print("generic sort ok", k.size, flush=True)
PY
for tool in synccheck racecheck; do
  timeout -k 10 400 compute-sanitizer --tool $tool --print-limit 5 python "$OUT/generic_sort.py" > "$OUT/sanitizer_$tool.log" 2>&1; echo "== $tool rc=$? (t+$(( $(date +%s) - T0 ))s)"; grep -E "ERROR SUMMARY|generic sort ok|Barrier error|hazard" "$OUT/sanitizer_$tool.log" | head -5
done
FL="--no-extra --no-cpu --no-nca --no-verify"
for mb in 0 4 6 8; do
  TFCUDA_MIN_BLOCKS=$mb timeout -k 10 200 python bench.py $FL > "$OUT/fluid_minblocks_$mb.log" 2> "$OUT/fluid_minblocks_$mb.err"
  python - "$OUT/fluid_minblocks_$mb.log" $mb <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    top = {r["name"]: round(r["total_ms"] / r["launches"] * 1e3, 1) for r in j["top_kernels"][:8]}
    print("min_blocks", sys.argv[2], "fluid ms", round(j["ms_per_step"], 4), top)
except Exception as e:
    print("min_blocks", sys.argv[2], "failed", e)
PY
done
cat > "$OUT/nbody_approx.py" <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
import tensorfrost_b200
from tensorfrost_b200 import workloads
tf = tensorfrost_b200.load()
nb = 262144
rng = np.random.default_rng(0)
hx = (5.0 * rng.standard_normal((nb, 3))).astype(np.float32)
x, v = tf.cuda_tensor(hx), tf.cuda_tensor(np.zeros((nb, 3), np.float32))
for name, fn in (("n_body", workloads.compile_nbody), ("n_body_loop", workloads.compile_nbody_loop)):
    prog = fn(tf)
    prog(x, v); tf.cuda_synchronize(); tf.cuda_timer_begin()
    for _ in range(2): out = prog(x, v)
    ms = tf.cuda_timer_end() / 2
    print(name, os.environ.get("TFCUDA_KERNEL_OPTIONS", ""), round(ms, 2), "ms", round(nb * nb / ms / 1e6, 1), "Ginteractions/s", flush=True)
PY
timeout -k 10 200 python "$OUT/nbody_approx.py" > "$OUT/nbody_ieee.log" 2>&1; grep Ginter "$OUT/nbody_ieee.log"
TFCUDA_KERNEL_OPTIONS="--prec-div=false --prec-sqrt=false" timeout -k 10 200 python "$OUT/nbody_approx.py" > "$OUT/nbody_approx.log" 2>&1; grep Ginter "$OUT/nbody_approx.log"
echo "total $(( $(date +%s) - T0 ))s"
