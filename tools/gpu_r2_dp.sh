#!/usr/bin/env bash
# N-GPU call: the peer-memory exchange against NCCL, then the driver's own scaling command at N.
#   gpurun --gpus 2 -- bash tools/gpu_r2_dp.sh 2
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
OUT=gpurun_out/r2dp$N; mkdir -p "$OUT"
nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29521 tests/dp_exchange_check.py \
    > "$OUT/exchange.log" 2>&1; echo "exchange rc=$?"; grep -E "exchange\]|EXCHANGE-OK|Error|error" "$OUT/exchange.log" | head -20
export TFCUDA_BENCH_DEADLINE=500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus "$N" --steps 8 --warmup 3 \
    > "$OUT/bench_n$N.json" 2> "$OUT/bench_n$N.err"; echo "bench N=$N rc=$?"; tail -c 3000 "$OUT/bench_n$N.json"; tail -5 "$OUT/bench_n$N.err"
TFCUDA_DP_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus "$N" --steps 8 --warmup 3 --no-single \
    > "$OUT/bench_n${N}_nccl.json" 2> "$OUT/bench_n${N}_nccl.err"; echo "bench nccl N=$N rc=$?"; cut -c1-400 "$OUT/bench_n${N}_nccl.json"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus "$N" --steps 10 --warmup 3 --workload fluid \
    > "$OUT/bench_fluid_replicas_n$N.json" 2> "$OUT/bench_fluid_replicas_n$N.err"; echo "fluid replicas N=$N rc=$?"; cut -c1-300 "$OUT/bench_fluid_replicas_n$N.json"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29525 bench.py --impl reference --gpus "$N" --steps 3 --warmup 1 \
    > "$OUT/bench_ref_n$N.json" 2> "$OUT/bench_ref_n$N.err"; echo "reference arm N=$N rc=$?"; cut -c1-300 "$OUT/bench_ref_n$N.json"
