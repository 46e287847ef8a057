#!/usr/bin/env bash
# Round 2, short 1-GPU call: is the per-GPU NCA step of the 8-GPU config (batch 32) host-bound?  graph on / off, device kernel total.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2e; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
run() {
  local name=$1 t=$2; shift 2
  local s=$(date +%s)
  timeout -k 10 "$t" stdbuf -oL -eL "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "== $name rc=$? $(( $(date +%s) - s ))s (t+$(( $(date +%s) - T0 ))s)"
}
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("ms_per_step", "gpu_launches", "host_issue_ms_per_step", "graph", "loss_after")})
    tk = j.get("top_kernels") or []
    for r in tk[:4]: print("    ", r)
    if tk: print("    ", tk[-1])
except Exception as e:
    print("no json:", e, open(sys.argv[1] + "".replace(".log", ".err")).read()[-500:] if False else "")
PY
}
NC="--workload nca --nca-batch 32 --nca-pool 128 --steps 20 --warmup 5 --nca-profile"
run nca_b32_default 300 python bench.py $NC; show "$OUT/nca_b32_default.log"; tail -2 "$OUT/nca_b32_default.err"
TFCUDA_GRAPH=0 run nca_b32_eager 300 python bench.py $NC; show "$OUT/nca_b32_eager.log"
run pytest_quick 600 python -m pytest tests/test_graph_replay_gpu.py tests/test_copy_engine_gpu.py tests/test_nca_gpu.py tests/test_parity_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "not live"; tail -4 "$OUT/pytest_quick.log"
run bench_fluid 300 python bench.py --no-extra --no-cpu --no-nca --no-verify; python - "$OUT/bench_fluid.log" <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fluid ms", j["ms_per_step"], "GB/s", j["value"], "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"], "graph", j["graph_replay"])
PY
echo "total $(( $(date +%s) - T0 ))s"
