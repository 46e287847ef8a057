#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=gpurun_out/r2dp${N}c; mkdir -p "$OUT"
export TFCUDA_BENCH_DEADLINE=300
TFCUDA_NCA_DIAG=1 timeout -k 10 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus "$N" --steps 20 --warmup 5 --no-single \
    > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench rc=$?"
python - "$OUT/bench_n$N.json" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("value", "ms_per_step", "graph", "diag_ms_per_step", "verify", "host_issue_ms_per_step", "per_rank_ms_per_step")})
except Exception as e:
    print("no json", e)
PY
tail -3 "$OUT/bench_n$N.err"
