#!/usr/bin/env bash
# NCA split-step bisect under the poison aid: which lowering reads memory it never wrote?
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TFCUDA_POISON=1
{
timeout 300 python tools/nca_debug.py 2>/dev/null | grep -A5 "^env="
TFCUDA_LIBRARY=0 timeout 300 python tools/nca_debug.py 2>/dev/null | grep -A5 "^env="
TFCUDA_LIBRARY_MATMUL=0 timeout 300 python tools/nca_debug.py 2>/dev/null | grep -A5 "^env="
TFCUDA_LIBRARY_MIN_AXIS=1000000000 timeout 300 python tools/nca_debug.py 2>/dev/null | grep -A5 "^env="
TFCUDA_RO=0 TFCUDA_DEFAULT_GROUP=0 TFCUDA_LIBRARY=0 timeout 300 python tools/nca_debug.py 2>/dev/null | grep -A5 "^env="
} > gpurun_out/nca_debug.txt 2>&1
cat gpurun_out/nca_debug.txt
