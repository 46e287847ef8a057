#!/usr/bin/env bash
# 8-GPU call (charged 8x): the exchange check, then exactly what the driver's scaling step runs at N = 8.
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=gpurun_out/r2dp$N; mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout -k 10 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29531 tests/dp_exchange_check.py \
    > "$OUT/exchange.log" 2>&1; echo "exchange rc=$?"; grep -E "exchange\]|EXCHANGE-OK|Error|error|Timeout" "$OUT/exchange.log" | head -12
export TFCUDA_BENCH_DEADLINE=420
timeout -k 10 480 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus "$N" --steps 20 --warmup 5 \
    > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench N=$N rc=$?"; tail -c 2600 "$OUT/bench_n$N.json"; tail -4 "$OUT/bench_n$N.err"
