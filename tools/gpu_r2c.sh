#!/usr/bin/env bash
# Round 2, 1-GPU call after the first measurements: re-validate, get the profiles that were lost (outputs must stay < 64 MiB).
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2c; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
run() {
  local name=$1 t=$2; shift 2
  local s=$(date +%s)
  timeout -k 10 "$t" stdbuf -oL -eL "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "== $name rc=$? $(( $(date +%s) - s ))s (t+$(( $(date +%s) - T0 ))s)"
}
run pytest_gpu 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --durations=8; tail -22 "$OUT/pytest_gpu.log"
run bench 1200 python bench.py; cut -c1-400 "$OUT/bench.log"; echo; python - "$OUT/bench.log" <<'PY'
import json, sys
j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"], "graph", j.get("graph_replay"))
print("roofline", {k: j["roofline"][k] for k in ("kernel", "achieved", "frac", "share_of_step", "launch_ms")}, j["roofline"]["whole_step"])
print("verify", j.get("verify"))
print("cpu_baseline", j.get("cpu_baseline"))
print("nca_dp", j.get("nca_dp"))
for k, v in j.get("extra_summary", {}).items(): print("  ", k, v)
for k, v in j.get("extra", {}).items():
    if isinstance(v, dict) and "verify" in v: print("  verify", k, v["verify"])
    if k.endswith("_error"): print("  ERROR", k, v)
for r in j["top_kernels"][:16]: print("  ", r["name"], r["launches"], round(r["total_ms"], 4), round(r["bytes"] / max(r["total_ms"], 1e-9) / 1e6, 1), "GB/s")
PY
NC="--workload nca --steps 5 --warmup 3 --nca-profile"
show() { python - "$1" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k: j.get(k) for k in ("value", "ms_per_step", "gpu_launches", "host_issue_ms_per_step", "device_alloc_calls_per_step", "graph", "loss_after", "build_seconds")})
    for r in (j.get("top_kernels") or [])[:14]: print("    ", r)
    print("    ", (j.get("top_kernels") or [None])[-1])
except Exception as e:
    print("no json:", e)
PY
}
run nca_default 400 python bench.py $NC; show "$OUT/nca_default.log"
TFCUDA_GRAPH=0 run nca_eager 400 python bench.py $NC; show "$OUT/nca_eager.log"
TFCUDA_MATMUL_MODE=0 run nca_tf32 400 python bench.py $NC; show "$OUT/nca_tf32.log"
FL="--no-extra --no-cpu --no-nca --no-verify"
run ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/fluid_launches.csv" python bench.py --steps 2 --warmup 3 $FL
run ncu_fluid_full 600 ncu --set full --clock-control none --import-source on -k 'regex:^kernel_(0|1|2|8|12|14)$' --launch-skip 60 -c 12 -f -o "$OUT/fluid_full" python bench.py --steps 2 --warmup 3 $FL
run ncu_fluid_csv 120 ncu -i "$OUT/fluid_full.ncu-rep" --page raw --csv; mv "$OUT/ncu_fluid_csv.log" "$OUT/fluid_full_raw.csv"
run ncu_sort_full 600 ncu --set full --clock-control none --import-source on -k 'regex:onesweep|digit_histogram' -c 6 -f -o "$OUT/sort_full" python tools/lib_kernels_once.py --medium
run ncu_sort_csv 120 ncu -i "$OUT/sort_full.ncu-rep" --page raw --csv; mv "$OUT/ncu_sort_csv.log" "$OUT/sort_full_raw.csv"
run ncu_lib_full 700 ncu --set full --clock-control none -k 'regex:gemm_tf32|reduce_rows|scan_rows|nbody_kernel|matmul_tn_kernel|scatter_add' -c 14 -f -o "$OUT/lib_full" python tools/lib_kernels_once.py --medium
run ncu_lib_csv 120 ncu -i "$OUT/lib_full.ncu-rep" --page raw --csv; mv "$OUT/ncu_lib_csv.log" "$OUT/lib_full_raw.csv"
rm -f "$OUT/lib_full.ncu-rep"
du -sh "$OUT"; ls -la "$OUT" | awk '{print $5, $9}' | sort -n | tail -8
echo "total $(( $(date +%s) - T0 ))s"
