#!/usr/bin/env bash
# One gpurun call for a development round: NCA bisect, full GPU test suite (no -x), smoke, bench, ncu launch list + full capture.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out; mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.txt" 2>&1
for what in "$@"; do
case $what in
ncadebug) bash tools/gpu_debug_nca.sh ;;
tests) timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > "$OUT/pytest_gpu.log" 2>&1; echo "pytest rc=$?"; tail -15 "$OUT/pytest_gpu.log" ;;
smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?"; tail -3 "$OUT/smoke.log" ;;
bench) timeout 1500 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench rc=$?"; tail -c 2500 "$OUT/bench.json"; tail -5 "$OUT/bench.err" ;;
benchfluid) timeout 600 python bench.py --no-extra --no-cpu > "$OUT/bench_fluid.json" 2> "$OUT/bench_fluid.err"; echo "bench rc=$?"; tail -c 2500 "$OUT/bench_fluid.json"; tail -5 "$OUT/bench_fluid.err" ;;
nculist) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_fluid.csv" \
      python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > "$OUT/ncu_launches.log" 2>&1; echo "ncu launches rc=$?" ;;
ncufull) timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:^kernel_(0|1|8|14)$' --launch-skip 45 -c 15 -f -o "$OUT/fluid_full" \
      python bench.py --steps 2 --warmup 3 --no-extra --no-cpu > "$OUT/ncu_full.log" 2>&1; echo "ncu full rc=$?"; ls -la "$OUT"/*.ncu-rep ;;
nca1) timeout 1500 python bench.py --workload nca --steps 5 --warmup 3 --nca-profile > "$OUT/nca_full_1.json" 2> "$OUT/nca_full_1.err"; echo "nca rc=$?"; tail -c 3000 "$OUT/nca_full_1.json"; tail -3 "$OUT/nca_full_1.err" ;;
esac
done
