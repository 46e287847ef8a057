#!/usr/bin/env bash
# 8-GPU call (charged 8x): where do the 10 ms between the per-rank program alone (64 ms) and the data-parallel step (74 ms) go?
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
OUT=gpurun_out/r2dp${N}b; mkdir -p "$OUT"
export TFCUDA_BENCH_DEADLINE=300
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("ms_per_step", "value", "graph", "diag_ms_per_step", "weak", "verify", "host_issue_ms_per_step")})
except Exception as e:
    print("no json:", e)
PY
}
TFCUDA_NCA_DIAG=1 timeout -k 10 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus "$N" --steps 20 --warmup 8 --no-single \
    > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"; echo "default rc=$?"; show "$OUT/bench_default.json"; tail -2 "$OUT/bench_default.err"
TFCUDA_GRAPH=0 timeout -k 10 360 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus "$N" --steps 20 --warmup 8 --no-single \
    > "$OUT/bench_eager.json" 2> "$OUT/bench_eager.err"; echo "eager rc=$?"; show "$OUT/bench_eager.json"
