#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out/r2dpmin$N; mkdir -p "$OUT"
export TFCUDA_BENCH_DEADLINE=100
timeout -k 5 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus "$N" --steps 6 --warmup 3 --no-single \
    > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench rc=$?"
python - "$OUT/bench_n$N.json" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("value", "ms_per_step", "warmup", "e2e", "verify", "per_rank_ms_per_step")})
except Exception as e:
    print("no json", e)
PY
tail -3 "$OUT/bench_n$N.err"
