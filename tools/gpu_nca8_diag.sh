#!/usr/bin/env bash
# 8-GPU box diagnosis: (A) 8 independent single-GPU NCA runs at the per-rank batch, concurrently, no torch / NCCL;
# (B) the data-parallel run.  Same per-GPU work in both.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
nproc > "$OUT/nca8_diag.txt"; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|^CPU\(s\)" >> "$OUT/nca8_diag.txt"
for i in 0 1 2 3 4 5 6 7; do
  CUDA_VISIBLE_DEVICES=$i timeout 500 python bench.py --workload nca --nca-batch 32 --nca-pool 128 --steps 5 --warmup 3 > "$OUT/nca_indep_$i.json" 2>/dev/null &
done
wait
for i in 0 1 2 3 4 5 6 7; do python -c "
import json; j=json.load(open('$OUT/nca_indep_$i.json')); print('independent gpu $i', round(j['ms_per_step'],1), 'ms host', round(j['host_issue_ms_per_step'],1))" >> "$OUT/nca8_diag.txt"; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload nca --steps 5 --warmup 3 > "$OUT/nca_full_8.json" 2> "$OUT/nca_full_8.err"
python -c "
import json; j=json.load(open('$OUT/nca_full_8.json')); print('dp8', round(j['ms_per_step'],1), j['per_rank_ms_per_step'], j['host_cores'])" >> "$OUT/nca8_diag.txt"
cat "$OUT/nca8_diag.txt"
