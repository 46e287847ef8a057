#!/usr/bin/env bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { echo "== $*"; "$@" 2>&1 | grep -E "passed|failed|AssertionError: grad" | tail -3; }
{
run python -m pytest tests/test_nca_gpu.py -q -x -k split_step
run python -m pytest tests/test_library_gpu.py tests/test_nca_gpu.py -q -k "matmul or split_step"
run python -m pytest tests/test_library_gpu.py tests/test_nca_gpu.py -q -k "reduce or split_step"
run python -m pytest tests/test_library_gpu.py tests/test_nca_gpu.py -q -k "sort or scatter or nbody or prefix or split_step"
run env TFCUDA_LIBRARY_MATMUL=0 python -m pytest tests/test_library_gpu.py tests/test_nca_gpu.py -q -k "matmul or split_step"
run env TFCUDA_LIBRARY=0 python -m pytest tests/test_library_gpu.py tests/test_nca_gpu.py -q -k "split_step or test_"
run python -m pytest tests/test_library_gpu.py tests/test_nca_gpu.py -q -k "matmul_ffma or split_step"
run python -m pytest tests/test_library_gpu.py tests/test_nca_gpu.py -q -k "matmul_tcgen05 or split_step"
} > gpurun_out/bisect.txt 2>&1
cat gpurun_out/bisect.txt
