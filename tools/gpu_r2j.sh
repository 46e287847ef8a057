#!/usr/bin/env bash
# Round 2, short 1-GPU call: the second-sight recorder policy (replay test, fluid, NCA per-GPU program).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2j; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
timeout -k 10 400 python -m pytest tests/test_graph_replay_gpu.py tests/test_copy_engine_gpu.py tests/test_zz_fluid_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > "$OUT/pytest.log" 2>&1; echo "pytest rc=$? (t+$(( $(date +%s) - T0 ))s)"; tail -3 "$OUT/pytest.log"
timeout -k 10 200 python bench.py --no-extra --no-cpu --no-nca --no-verify > "$OUT/fluid.log" 2> "$OUT/fluid.err"; python - "$OUT/fluid.log" <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("fluid ms", j["ms_per_step"], "GB/s", j["value"], "e2e ms", j["e2e"]["ms_per_step"], "graph", j["graph_replay"])
PY
timeout -k 10 300 python bench.py --workload nca --nca-batch 32 --nca-pool 128 --steps 20 --warmup 5 > "$OUT/nca_b32.log" 2> "$OUT/nca_b32.err"; python - "$OUT/nca_b32.log" <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("nca b32 ms", j["ms_per_step"], "graph", j["graph"], "host_issue", j["host_issue_ms_per_step"])
PY
echo "total $(( $(date +%s) - T0 ))s"
