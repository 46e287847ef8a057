#!/usr/bin/env bash
# Builds the CUDA-enabled TensorFrost python module into build/tf_cuda/TensorFrost/.
#
# = reference sources (scratch copy, never stored in git) + the insertions of overlay/apply_overlay.py
#   + our backend glue and CUDA emitter (overlay/Backend/**), linked against tensorfrost_b200/lib/libtfcuda.so.
# The reference's frontend, IR and compiler passes are used unchanged; user code switches backend with
# tf.initialize(tf.cuda).  On the GPU box /root/reference is absent and the prebuilt module is used.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
REPO=$(cd "$HERE/.." && pwd)
REF=${TF_REFERENCE:-/root/reference}
OUT=$REPO/build/tf_cuda
SCRATCH=${TF_OVERLAY_SCRATCH:-/tmp/tf_overlay_build}
if [ ! -d "$REF/TensorFrost" ]; then
  echo "[overlay] $REF not present; using prebuilt $OUT if any" >&2
  exit 0
fi
[ -f "$HERE/lib/libtfcuda.so" ] || "$HERE/build_lib.sh"
REFHASH=$( (cd "$REF" && find TensorFrost Python CMakeLists.txt -type f \( -name '*.cpp' -o -name '*.h' -o -name '*.py' -o -name 'CMakeLists.txt' \) -print0 | sort -z | xargs -0 sha1sum; sha1sum "$HERE/overlay/apply_overlay.py") | sha1sum | cut -d' ' -f1)
S=$SCRATCH/src
if [ ! -f "$SCRATCH/.refhash" ] || [ "$(cat "$SCRATCH/.refhash")" != "$REFHASH" ]; then
  rm -rf "$SCRATCH"; mkdir -p "$SCRATCH"
  cp -r "$REF" "$S"; chmod -R u+w "$S"
  python3 "$HERE/overlay/apply_overlay.py" "$S" "$REPO"
  cmake -S "$S" -B "$SCRATCH/build" -G Ninja -DCMAKE_BUILD_TYPE=Release \
    -DGLFW_BUILD_X11=OFF -DGLFW_BUILD_WAYLAND=OFF -DCMAKE_POSITION_INDEPENDENT_CODE=ON \
    --compile-no-warning-as-error \
    -DCMAKE_CXX_FLAGS="-DGLFW_INCLUDE_NONE -fkeep-inline-functions" > "$SCRATCH/cmake.log" 2>&1
  echo "$REFHASH" > "$SCRATCH/.refhash"
else
  # incremental: refresh only our own sources
  for rel in Backend/Backends/CUDA/CUDA.h Backend/Backends/CUDA/CudaBackend.cpp Backend/Backends/CUDA/CudaLibrary.cpp Backend/Backends/CUDA/CudaPython.cpp Backend/CodeGen/Langs/CUDA.cpp; do
    cmp -s "$HERE/overlay/$rel" "$S/TensorFrost/$rel" || cp "$HERE/overlay/$rel" "$S/TensorFrost/$rel"
  done
fi
if ! ninja -C "$SCRATCH/build" TensorFrost > "$SCRATCH/ninja.log" 2>&1; then
  grep -E "error|Error" -A5 "$SCRATCH/ninja.log" | head -80 >&2
  exit 1
fi
rm -rf "$OUT/TensorFrost"; mkdir -p "$OUT/TensorFrost"
cp "$S"/Python/TensorFrost/*.py "$OUT/TensorFrost/"
cp "$S"/Python/TensorFrost/TensorFrost*.so "$OUT/TensorFrost/"
# the module's RPATH is $ORIGIN: keep the runtime library next to it
cp "$HERE/lib/libtfcuda.so" "$OUT/TensorFrost/libtfcuda.so"
# backend-aware python helpers of ours that extend the package (e.g. the sort dispatch)
if [ -d "$HERE/overlay/python" ]; then
  python3 "$HERE/overlay/python/install.py" "$OUT/TensorFrost"
fi
echo "[overlay] built $(ls "$OUT"/TensorFrost/TensorFrost*.so)"
