#!/usr/bin/env bash
# N-GPU call, short: the exchange check and the driver's scaling command at N.
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out/r2dps$N; mkdir -p "$OUT"
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29551 tests/dp_exchange_check.py \
    > "$OUT/exchange.log" 2>&1; echo "exchange rc=$?"; grep -E "exchange\]|EXCHANGE-OK|Error|error|Timeout" "$OUT/exchange.log" | head -8
export TFCUDA_BENCH_DEADLINE=420
timeout -k 10 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus "$N" --steps 20 --warmup 5 \
    > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench N=$N rc=$?"
python - "$OUT/bench_n$N.json" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("metric", "value", "ms_per_step", "n_gpus", "graph", "weak", "single_gpu", "verify")})
    print(j["config"])
except Exception as e:
    print("no json", e)
PY
tail -3 "$OUT/bench_n$N.err"
