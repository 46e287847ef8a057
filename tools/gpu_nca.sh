#!/usr/bin/env bash
# NCA on the GPU box: parity tests, then bench at a small and at the BASELINE config.  usage: bash tools/gpu_nca.sh [ngpus]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
N=${1:-1}
if [ "$N" = 1 ]; then
  timeout 900 python -m pytest tests/test_nca_gpu.py -q -x --timeout 800 > "$OUT/pytest_nca.log" 2>&1; echo "pytest nca rc=$?"; tail -15 "$OUT/pytest_nca.log"
  timeout 900 python bench.py --workload nca --nca-batch 32 --nca-grid 64 --steps 5 --warmup 3 > "$OUT/nca_small.json" 2> "$OUT/nca_small.err"; echo "small rc=$?"; cat "$OUT/nca_small.json"; tail -3 "$OUT/nca_small.err"
  timeout 1500 python bench.py --workload nca --steps 5 --warmup 3 --nca-profile > "$OUT/nca_full_1.json" 2> "$OUT/nca_full_1.err"; echo "full rc=$?"; cat "$OUT/nca_full_1.json"; tail -3 "$OUT/nca_full_1.err"
  nvidia-smi --query-gpu=memory.used,memory.total --format=csv
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$N" --workload nca --steps 5 --warmup 3 \
      > "$OUT/nca_full_$N.json" 2> "$OUT/nca_full_$N.err"; echo "dp$N rc=$?"; cat "$OUT/nca_full_$N.json"; tail -5 "$OUT/nca_full_$N.err"
fi
