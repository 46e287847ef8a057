#!/usr/bin/env bash
# Round 2, validation of the final build on 1 GPU: what the driver runs at round end (tests, smoke, bench), outputs kept small.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2f; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
run() {
  local name=$1 t=$2; shift 2
  local s=$(date +%s)
  timeout -k 10 "$t" stdbuf -oL -eL "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "== $name rc=$? $(( $(date +%s) - s ))s (t+$(( $(date +%s) - T0 ))s)"
}
run pytest_gpu 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=5; tail -12 "$OUT/pytest_gpu.log"
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"; tail -1 "$OUT/smoke.log"
run bench 1200 python bench.py --gpus 1 --steps 20 --warmup 5; python - "$OUT/bench.log" <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["ms_per_step"], j["e2e"]["value"], "graph", j.get("graph_replay"))
print("roofline", {k: j["roofline"][k] for k in ("kernel", "achieved", "frac", "share_of_step", "launch_ms", "traffic")}, j["roofline"]["whole_step"])
print("verify", (j.get("verify") or {}).get("ok"), "cpu", j.get("cpu_baseline", {}).get("value"), "nca_dp", {k: (j.get("nca_dp") or {}).get(k) for k in ("value", "ms_per_step", "error")})
for k, v in j.get("extra_summary", {}).items(): print("  ", k, v)
for k, v in j.get("extra", {}).items():
    if k.endswith("_error"): print("  ERROR", k, v)
PY
run bench_ref 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5; cut -c1-500 "$OUT/bench_ref.log"
echo "total $(( $(date +%s) - T0 ))s"; du -sh "$OUT"
