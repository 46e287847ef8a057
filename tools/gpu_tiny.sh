#!/usr/bin/env bash
# The last ~35 seconds of the round's GPU budget: the library matmuls (new epilogue store, new matmul_tn layout) against float64,
# then one short fluid run with the lane code.  Every step under its own timeout.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/tiny; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
timeout -k 3 17 python -m pytest tests/test_library_gpu.py -k "matmul" -x -q -p no:cacheprovider > "$OUT/pytest_matmul.log" 2>&1; echo "matmul rc=$? (t+$(( $(date +%s) - T0 ))s)"; tail -4 "$OUT/pytest_matmul.log"
timeout -k 3 16 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu --no-nca --no-verify > "$OUT/fluid.log" 2> "$OUT/fluid.err"; echo "fluid rc=$? (t+$(( $(date +%s) - T0 ))s)"
python - "$OUT/fluid.log" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("fluid ms/step", j["ms_per_step"], "GB/s", j["value"], "emitter", j.get("emitter"), [(k["name"], round(k["total_ms"] / k["launches"] * 1000, 1)) for k in j["top_kernels"]])
except Exception as e:
    print("no fluid line:", e)
PY
