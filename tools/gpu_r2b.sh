#!/usr/bin/env bash
# Round 2 main 1-GPU call: staged (cheap checks first), every stage under `timeout -k`, unbuffered logs.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2b; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
run() {  # name, timeout, command...
  local name=$1 t=$2; shift 2
  local s=$(date +%s)
  timeout -k 10 "$t" stdbuf -oL -eL "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "== $name rc=$? $(( $(date +%s) - s ))s (t+$(( $(date +%s) - T0 ))s)"
}
cat > "$OUT/step.py" <<'PY'
import faulthandler, sys, os
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import tensorfrost_b200
tf = tensorfrost_b200.load()
import cases
for name in sys.argv[1:]:
    for rep in range(3):
        outs, _ = cases.run_case(tf, name, seed=0)
    print("case", name, "ok", flush=True)
print("stats", tf.cuda_graph_stats(), flush=True)
tf.cuda_synchronize()
print("DONE", flush=True)
PY
run step_basic 120 python "$OUT/step.py" wave host_loop atomics sort_radix_u32 matmul; tail -2 "$OUT/step_basic.log"
run smoke 200 python -c "import __graft_entry__ as g; g.smoke()"; tail -1 "$OUT/smoke.log"
run pytest_gpu 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --durations=15; tail -30 "$OUT/pytest_gpu.log"
run bench 1200 python bench.py; cut -c1-1200 "$OUT/bench.log"; echo; tail -c 3500 "$OUT/bench.log"; tail -3 "$OUT/bench.err"
FL="--no-extra --no-cpu --no-nca --no-verify"
TFCUDA_GRAPH=0 run bench_eager 300 python bench.py $FL; cut -c1-330 "$OUT/bench_eager.log"; echo
TFCUDA_PDL=1 run bench_pdl_graph 300 python bench.py $FL; cut -c1-330 "$OUT/bench_pdl_graph.log"; echo
TFCUDA_KERNEL_OPTIONS="--prec-div=false --prec-sqrt=false" run bench_approx 400 python bench.py --no-extra --no-cpu --no-nca; cut -c1-330 "$OUT/bench_approx.log"; echo; grep -o '"verify": {[^}]*}[^}]*}' "$OUT/bench_approx.log" | head -2
TFCUDA_KERNEL_OPTIONS="--prec-div=false --prec-sqrt=false" run pytest_approx 600 python -m pytest tests/test_parity_gpu.py tests/test_zz_fluid_gpu.py tests/test_nca_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "not live"; tail -8 "$OUT/pytest_approx.log"
# NCA 1 GPU: default (3xTF32), single TF32, resident-weight FFMA kernel, plain atomics
NC="--workload nca --steps 5 --warmup 3 --nca-profile"
run nca_default 400 python bench.py $NC; cut -c1-260 "$OUT/nca_default.log"; echo
TFCUDA_MATMUL_MODE=0 run nca_tf32 400 python bench.py $NC; cut -c1-260 "$OUT/nca_tf32.log"; echo
TFCUDA_MATMUL_ROWS=1 run nca_rows 400 python bench.py $NC; cut -c1-260 "$OUT/nca_rows.log"; echo
TFCUDA_MATMUL_MODE=0 TFCUDA_KERNEL_OPTIONS="-DTF_WARP_AGG_ATOMICS=0" run nca_tf32_plain_atomics 400 python bench.py $NC; cut -c1-260 "$OUT/nca_tf32_plain_atomics.log"; echo
TFCUDA_MATMUL_MODE=0 run pytest_nca_tf32 400 python -m pytest tests/test_nca_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider; tail -3 "$OUT/pytest_nca_tf32.log"
# ncu: launch list of the fluid step (same command as the bench line), full captures of its kernel classes and of the sort
run ncu_launches 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/fluid_launches.csv" python bench.py --steps 2 --warmup 3 $FL
run ncu_fluid_full 600 ncu --set full --clock-control none --import-source on -k 'regex:^kernel_(0|1|2|8|12|14)$' --launch-skip 60 -c 12 -f -o "$OUT/fluid_full" python bench.py --steps 2 --warmup 3 $FL
run ncu_lib_full 700 ncu --set full --clock-control none --import-source on -k 'regex:onesweep|digit_histogram|gemm_tf32|reduce_rows|scan_rows|nbody|matmul_tn_kernel|matmul_rows' -c 24 -f -o "$OUT/lib_full" python tools/lib_kernels_once.py --medium
ls -la "$OUT" | head -60
echo "total $(( $(date +%s) - T0 ))s"
