#!/usr/bin/env bash
# Round 2, short 1-GPU call: whole-row blocks in NCA (A/B), e2e after the host-callback fix, NCA parity.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2h; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
run() {
  local name=$1 t=$2; shift 2
  local s=$(date +%s)
  timeout -k 10 "$t" stdbuf -oL -eL "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "== $name rc=$? $(( $(date +%s) - s ))s (t+$(( $(date +%s) - T0 ))s)"
}
show() { python - "$1" <<'PY'
import json, sys, collections
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("ms_per_step", "gpu_launches", "loss_after")})
    tk = j.get("top_kernels") or []
    for r in tk[:3]: print("    ", r)
    if tk: print("    ", tk[-1])
except Exception as e:
    print("no json:", e)
PY
}
NC="--workload nca --steps 5 --warmup 3 --nca-profile"
export TFCUDA_PROFILE_DUMP="$OUT/nca_profile_whole_rows.json"
run nca_whole_rows 400 python bench.py $NC; show "$OUT/nca_whole_rows.log"
export TFCUDA_PROFILE_DUMP="$OUT/nca_profile_32wide.json"
TFCUDA_WHOLE_ROWS=0 run nca_32wide 400 python bench.py $NC; show "$OUT/nca_32wide.log"
unset TFCUDA_PROFILE_DUMP
run pytest_nca 600 python -m pytest tests/test_nca_gpu.py tests/test_copy_engine_gpu.py tests/test_interop_gpu.py tests/test_op_table_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider; tail -3 "$OUT/pytest_nca.log"
run bench_fluid 300 python bench.py --no-extra --no-cpu --no-nca --no-verify; python - "$OUT/bench_fluid.log" <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fluid ms", j["ms_per_step"], "GB/s", j["value"], "e2e ms", j["e2e"]["ms_per_step"], "e2e GB/s", j["e2e"]["value"])
PY
echo "total $(( $(date +%s) - T0 ))s"
