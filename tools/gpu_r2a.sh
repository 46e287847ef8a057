#!/usr/bin/env bash
# Round 2, first GPU call (1 GPU): validate everything written since the last hardware run, then measure.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2a; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
nproc >> "$OUT/gpu.txt"; free -g | head -2 >> "$OUT/gpu.txt"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -2 "$OUT/smoke.log"
timeout 3000 python -m pytest tests -m gpu -q --timeout 1500 -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -25 "$OUT/pytest_gpu.log"
timeout 1500 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; cut -c1-1500 "$OUT/bench.json"; tail -c 2500 "$OUT/bench.json"; tail -5 "$OUT/bench.err"
# A/B of the launch path on the fluid step: eager launches, programmatic dependent launch
TFCUDA_GRAPH=0 timeout 600 python bench.py --no-extra --no-cpu --no-nca --no-verify > "$OUT/bench_eager.json" 2> "$OUT/bench_eager.err"; echo "eager rc=$?"; cut -c1-400 "$OUT/bench_eager.json"
TFCUDA_PDL=1 timeout 600 python bench.py --no-extra --no-cpu --no-nca --no-verify > "$OUT/bench_pdl.json" 2> "$OUT/bench_pdl.err"; echo "pdl rc=$?"; cut -c1-400 "$OUT/bench_pdl.json"
# skinny matmul inside the NCA programs: parity, then timing (with per-kernel profile)
TFCUDA_MATMUL_ROWS=1 timeout 900 python -m pytest tests/test_nca_gpu.py -m gpu -q -p no:cacheprovider > "$OUT/pytest_rows_nca.log" 2>&1; echo "rows nca parity rc=$?"; tail -3 "$OUT/pytest_rows_nca.log"
timeout 600 python bench.py --workload nca --steps 5 --warmup 3 --nca-profile > "$OUT/nca_1.json" 2> "$OUT/nca_1.err"; echo "nca rc=$?"; cut -c1-300 "$OUT/nca_1.json"
TFCUDA_MATMUL_ROWS=1 timeout 600 python bench.py --workload nca --steps 5 --warmup 3 --nca-profile > "$OUT/nca_1_rows.json" 2> "$OUT/nca_1_rows.err"; echo "nca rows rc=$?"; cut -c1-300 "$OUT/nca_1_rows.json"
# ncu: launch list of the fluid step (same command as the bench line), then full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/fluid_launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-nca --no-verify > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^kernel_(0|1|2|8|12|14)$' --launch-skip 60 -c 12 -f -o "$OUT/fluid_full" \
    python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-nca --no-verify > "$OUT/ncu_fluid_full.log" 2>&1; echo "ncu fluid rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:onesweep|digit_histogram|gemm_tf32|reduce_rows|reduce_mid|scan_rows|nbody|matmul_tn_kernel|matmul_rows|split_tf32|transpose_kernel' \
    -c 30 -f -o "$OUT/lib_full" python tools/lib_kernels_once.py > "$OUT/ncu_lib.log" 2>&1; echo "ncu lib rc=$?"
ls -la "$OUT"
