#!/usr/bin/env bash
# Round 2, short 1-GPU call: the ballot-ranked sort (tests + timing + ncu summary) and a probe of host<->device copy overlap.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2d; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1 PYTHONFAULTHANDLER=1
T0=$(date +%s)
run() {
  local name=$1 t=$2; shift 2
  local s=$(date +%s)
  timeout -k 10 "$t" stdbuf -oL -eL "$@" > "$OUT/$name.log" 2> "$OUT/$name.err"
  echo "== $name rc=$? $(( $(date +%s) - s ))s (t+$(( $(date +%s) - T0 ))s)"
}
run pytest_sort 600 python -m pytest tests/test_library_gpu.py tests/test_parity_gpu.py tests/test_zz_scale_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "sort or radix"; tail -4 "$OUT/pytest_sort.log"
cat > "$OUT/probe.py" <<'PY'
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
import tensorfrost_b200
tf = tensorfrost_b200.load()
rng = np.random.default_rng(0)
for logn in (26, 28):
    n = 1 << logn
    keys = tf.cuda_tensor(rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))
    vals = tf.cuda_tensor(np.arange(n, dtype=np.uint32))
    for name, fn in (("keys", lambda: tf.cuda_radix_sort(keys)), ("pairs", lambda: tf.cuda_radix_sort(keys, vals))):
        for _ in range(3): fn()
        tf.cuda_synchronize(); tf.cuda_timer_begin()
        for _ in range(5): fn()
        ms = tf.cuda_timer_end() / 5
        print(f"sort 2^{logn} {name}: {ms:.3f} ms = {n / ms / 1e6:.2f} Gkeys/s", flush=True)
    del keys, vals
# copy overlap probe: 64 MB up, 64 MB down, alone and together (page-locked host memory)
m = 16 << 20
a = tf.cuda_pinned_array([m], "float32"); b = tf.cuda_pinned_array([m], "float32")
a[...] = 1.0
d1 = tf.cuda_tensor(np.zeros(m, np.float32)); d2 = tf.cuda_tensor(np.ones(m, np.float32))
def timed(fn, reps=10):
    fn(); tf.cuda_copy_sync(); tf.cuda_synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    tf.cuda_copy_sync(); tf.cuda_synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
up = timed(lambda: tf.cuda_upload_async(d1, a))
down = timed(lambda: tf.cuda_download_async(d2, b))
both = timed(lambda: (tf.cuda_upload_async(d1, a), tf.cuda_download_async(d2, b)))
blocking = timed(lambda: (tf.cuda_upload(d1, a), tf.cuda_download(d2, b)))
print(f"copy 64 MB: upload {up:.3f} ms ({m * 4 / up / 1e6:.1f} GB/s), download {down:.3f} ms ({m * 4 / down / 1e6:.1f} GB/s), both async {both:.3f} ms, both blocking {blocking:.3f} ms", flush=True)
PY
run probe 300 python "$OUT/probe.py"; grep -E "^sort|^copy" "$OUT/probe.log"
run ncu_sort 400 ncu --set full --clock-control none -k 'regex:onesweep' -c 3 -f -o "$OUT/sort_full" python tools/lib_kernels_once.py --medium
run ncu_sort_csv 60 ncu -i "$OUT/sort_full.ncu-rep" --page raw --csv; mv "$OUT/ncu_sort_csv.log" "$OUT/sort_full_raw.csv"; rm -f "$OUT/sort_full.ncu-rep"
python - "$OUT/sort_full_raw.csv" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[i], rows[i + 2:]
col = {h: k for k, h in enumerate(hdr)}
for r in data:
    if len(r) >= len(hdr):
        print({k.split(".")[0]: r[col[k]] for k in ("gpu__time_duration.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
               "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum") if k in col})
PY
echo "total $(( $(date +%s) - T0 ))s"
