#!/usr/bin/env bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ended, in one box.
#   1 GPU : bash tools/gpu_next.sh            (tests incl. experimental, bench, PDL A/B, matmul_rows A/B on NCA)
#   8 GPUs: bash tools/gpu_next.sh dp 8       (NCA data parallel with the bracketed exchange; never pass --nca-profile with N > 1 unprotected)
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
if [ "${1:-}" = dp ]; then
  N=${2:-8}
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus "$N" --workload nca --steps 5 --warmup 3 \
      > "$OUT/nca_dp$N.json" 2> "$OUT/nca_dp$N.err"; echo "dp$N rc=$?"; cut -c1-400 "$OUT/nca_dp$N.json"
  TFCUDA_DP_SYNC=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus "$N" --workload nca --steps 5 --warmup 3 \
      > "$OUT/nca_dp${N}_async.json" 2> "$OUT/nca_dp${N}_async.err"; echo "dp$N async rc=$?"; cut -c1-400 "$OUT/nca_dp${N}_async.json"
  exit 0
fi
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -8 "$OUT/pytest_gpu.log"
TFCUDA_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_library_gpu.py -m gpu -q -k matmul_rows > "$OUT/pytest_experimental.log" 2>&1; echo "experimental rc=$?"; tail -5 "$OUT/pytest_experimental.log"
timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; cut -c1-600 "$OUT/bench.json"
# programmatic dependent launch: parity of the fluid program first, then the timing A/B
TFCUDA_PDL=1 timeout 600 python -m pytest tests/test_zz_fluid_gpu.py tests/test_parity_gpu.py -m gpu -q -k "fluid or wave or host_loop" > "$OUT/pytest_pdl.log" 2>&1; echo "pdl parity rc=$?"; tail -3 "$OUT/pytest_pdl.log"
TFCUDA_PDL=1 timeout 600 python bench.py --no-extra --no-cpu > "$OUT/bench_pdl.json" 2> "$OUT/bench_pdl.err"; echo "bench pdl rc=$?"; cut -c1-300 "$OUT/bench_pdl.json"
# skinny matmul inside the NCA programs: parity, then timing
TFCUDA_MATMUL_ROWS=1 timeout 600 python -m pytest tests/test_nca_gpu.py -m gpu -q > "$OUT/pytest_rows_nca.log" 2>&1; echo "rows nca parity rc=$?"; tail -3 "$OUT/pytest_rows_nca.log"
timeout 400 python bench.py --workload nca --steps 5 --warmup 3 > "$OUT/nca_1.json" 2>/dev/null; cut -c1-300 "$OUT/nca_1.json"
TFCUDA_MATMUL_ROWS=1 timeout 400 python bench.py --workload nca --steps 5 --warmup 3 > "$OUT/nca_1_rows.json" 2>/dev/null; cut -c1-300 "$OUT/nca_1_rows.json"
# ncu evidence for the library kernels (round 1 has it only for the emitted fluid kernels)
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:onesweep|digit_histogram|gemm_tf32|reduce_rows|reduce_mid|scan|nbody|matmul_tn|split_tf32|transpose' \
    -c 40 -f -o "$OUT/lib_full" python tools/lib_kernels_once.py > "$OUT/ncu_lib.log" 2>&1; echo "ncu lib rc=$?"
