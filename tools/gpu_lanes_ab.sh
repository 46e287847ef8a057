#!/usr/bin/env bash
# FIRST GPU call after the round-2 emitter / library work that was done without a device (DESIGN.md 3 "done after the GPU budget was
# spent"): the GPU suite, then A/B timings of every switch, then ncu of the kernels that changed.  One GPU, ~12 minutes.
#   gpurun --timeout 900 -- bash tools/gpu_lanes_ab.sh
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/lanes_ab; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
timeout -k 10 500 python -m pytest tests -m gpu -x -q --timeout 400 -p no:cacheprovider > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$? (t+$(( $(date +%s) - T0 ))s)"; tail -3 "$OUT/pytest_gpu.log"
summ() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(j["ms_per_step"], 4), "value", round(j["value"], 1), "emitter", j.get("emitter"),
          "top", [(k["name"], round(k["total_ms"] / k["launches"] * 1000, 1)) for k in (j.get("top_kernels") or [])[:8]])
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
}
# fluid 2048^2: default, no lanes, no range fact either (= the build measured in DESIGN.md 6), 2 lanes, lanes from 2^22 elements on only
for tag in default:"" nolanes:"TFCUDA_COARSEN=0" measured:"TFCUDA_COARSEN=0 TFCUDA_ASSUME=0" lanes2:"TFCUDA_COARSEN=2" from4m:"TFCUDA_COARSEN_MIN_ELEMENTS=4194304"; do
  name=${tag%%:*}; envs=${tag#*:}
  env $envs timeout -k 10 120 python bench.py --no-extra --no-cpu --no-nca --no-verify > "$OUT/fluid_$name.log" 2> "$OUT/fluid_$name.err"; summ "$OUT/fluid_$name.log"
done
# NCA, the per-GPU program of the 8-GPU config and the whole batch: default vs the measured build
for b in 32:128 256:1024; do
  batch=${b%%:*}; pool=${b#*:}
  for tag in default:"" measured:"TFCUDA_COARSEN=0 TFCUDA_ASSUME=0"; do
    name=${tag%%:*}; envs=${tag#*:}
    env $envs timeout -k 10 300 python bench.py --workload nca --nca-batch $batch --nca-pool $pool --steps 6 --warmup 12 --nca-profile > "$OUT/nca_b${batch}_$name.log" 2> "$OUT/nca_b${batch}_$name.err"
    summ "$OUT/nca_b${batch}_$name.log"
  done
done
# ncu: launch list of the fluid step, full capture of a coarsened stencil kernel, the projection kernel and the advection kernel
timeout -k 10 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/fluid_launches.csv" \
  python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-nca --no-verify > "$OUT/ncu_launches.log" 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k 'regex:^kernel_(0|2|12|14)$' --launch-skip 60 -c 8 -o "$OUT/fluid_full" -f \
  python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-nca --no-verify > "$OUT/ncu_full.log" 2>&1
ncu -i "$OUT/fluid_full.ncu-rep" --page raw --csv > "$OUT/fluid_full_raw.csv" 2>/dev/null
# the skinny products of NCA (epilogue store pattern) and the weight-gradient kernel at NCA's shapes
timeout -k 10 300 ncu --set full --clock-control none -k 'regex:gemm_tf32_kernel|matmul_tn_kernel' --launch-skip 40 -c 8 -o "$OUT/nca_matmul_full" -f \
  python bench.py --workload nca --nca-batch 32 --nca-pool 128 --steps 1 --warmup 1 > "$OUT/ncu_nca.log" 2>&1
ncu -i "$OUT/nca_matmul_full.ncu-rep" --page raw --csv > "$OUT/nca_matmul_full_raw.csv" 2>/dev/null
echo "total $(( $(date +%s) - T0 ))s"
