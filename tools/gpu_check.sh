#!/usr/bin/env bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list of the bench command and one full ncu
# capture of the dominant kernel.  Everything lands in gpurun_out/.  Each step has its own timeout so a hung
# kernel cannot hold the box.
#   usage (under gpurun): bash tools/gpu_check.sh [tests|bench|ncu|all] [pytest -k expression]
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p "$OUT"
WHAT=${1:-all}
KEXPR=${2:-}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
python -c "import os; print('cores', os.cpu_count())" >> "$OUT/gpu.txt"

if [ "$WHAT" = tests ] || [ "$WHAT" = all ]; then
  timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 ${KEXPR:+-k "$KEXPR"} > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest rc=$?" | tee -a "$OUT/pytest_gpu.log"
  tail -25 "$OUT/pytest_gpu.log"
fi
if [ "$WHAT" = smoke ] || [ "$WHAT" = all ]; then
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
  echo "smoke rc=$?" | tee -a "$OUT/smoke.log"
  tail -3 "$OUT/smoke.log"
fi
if [ "$WHAT" = bench ] || [ "$WHAT" = all ]; then
  timeout 1500 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"
  echo "bench rc=$?"
  tail -c 3000 "$OUT/bench.json"; tail -5 "$OUT/bench.err"
fi
if [ "$WHAT" = ncu ] || [ "$WHAT" = all ]; then
  # launch list of the bench command (fluid only): shares of the step per kernel
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_fluid.csv" \
      python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > "$OUT/ncu_launches.log" 2>&1
  echo "ncu launches rc=$?"
fi
