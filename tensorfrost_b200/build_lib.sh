#!/usr/bin/env bash
# Builds tensorfrost_b200/lib/libtfcuda.so for sm_100a with nvcc (cross-compiles without a GPU).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/csrc
OUT=$HERE/lib
OBJ=$HERE/../build/tfcuda_obj
mkdir -p "$OUT" "$OBJ"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -cudart static -I "$HERE/../include")
# embed the device prelude as a C string
python3 - "$SRC/prelude.cuh" "$SRC/prelude_embed.inc" <<'PY'
import sys
src = open(sys.argv[1]).read()
out = []
for line in src.splitlines():
    out.append('"' + line.replace('\\', '\\\\').replace('"', '\\"') + '\\n"')
text = "\n".join(out) + "\n"
try:
    old = open(sys.argv[2]).read()
except FileNotFoundError:
    old = None
if old != text:
    open(sys.argv[2], "w").write(text)
PY
objs=()
pids=()
for f in runtime comm reduce scan radix_sort scatter nbody matmul matmul_tn matmul_rows matmul_tcgen05; do
  o=$OBJ/$f.o
  objs+=("$o")
  if [ ! -f "$o" ] || [ "$SRC/$f.cu" -nt "$o" ] || [ "$SRC/tfcuda_internal.h" -nt "$o" ] || [ "$HERE/../include/tfcuda.h" -nt "$o" ] \
     || { [ "$f" = runtime ] && [ "$SRC/prelude_embed.inc" -nt "$o" ]; }; then
    ( "$NVCC" "${FLAGS[@]}" ${TFCUDA_PTXAS_V:+-Xptxas -v} -c "$SRC/$f.cu" -o "$o" ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -cudart static -o "$OUT/libtfcuda.so" "${objs[@]}" -L/usr/local/cuda/lib64 -lnvrtc -ldl -lpthread \
  -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
# the CUDA-enabled module loads the copy next to itself (RPATH $ORIGIN): keep it identical
if [ -d "$HERE/../build/tf_cuda/TensorFrost" ]; then cp "$OUT/libtfcuda.so" "$HERE/../build/tf_cuda/TensorFrost/libtfcuda.so"; fi
echo "[tfcuda] built $OUT/libtfcuda.so"
