#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2g; mkdir -p "$OUT"
export PYTHONUNBUFFERED=1
timeout -k 10 200 python tools/copy_overlap_probe.py > "$OUT/probe_default.log" 2>&1; grep -E "ms/step|HOSTFUNC" "$OUT/probe_default.log"
TFCUDA_COPY_HOSTFUNC=0 timeout -k 10 200 python tools/copy_overlap_probe.py > "$OUT/probe_nohostfunc.log" 2>&1; grep -E "ms/step|HOSTFUNC" "$OUT/probe_nohostfunc.log"
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout -k 10 200 python tools/copy_overlap_probe.py > "$OUT/probe_conn32.log" 2>&1; echo "CUDA_DEVICE_MAX_CONNECTIONS=32"; grep -E "ms/step" "$OUT/probe_conn32.log"
