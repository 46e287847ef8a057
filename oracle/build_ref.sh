#!/usr/bin/env bash
# Builds the UNMODIFIED reference (MichaelMoroz/TensorFrost) python module from the
# sources under /root/reference into oracle/_ref/TensorFrost/ (git-ignored, ships to the
# GPU box with gpurun).  The reference tree is read-only and its CMake writes into its own
# source dir, so we build from a scratch copy under $TMPDIR; nothing from the reference is
# stored in git.  Recipe: SURVEY.md §8(c).
#
# TEST INFRASTRUCTURE ONLY: the result is the parity oracle (tf.cpu, the reference's own
# C++/OpenMP backend) and bench.py's cpu_baseline / --impl reference arm.
set -euo pipefail
REF=${TF_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
SCRATCH=${TF_ORACLE_SCRATCH:-/tmp/tf_oracle_build}
if [ ! -d "$REF/TensorFrost" ]; then
  echo "[oracle] $REF not present; using prebuilt $OUT if any" >&2
  exit 0
fi
STAMP=$OUT/.stamp
WANT=$(cd "$REF" && find TensorFrost Python CMakeLists.txt -type f \( -name '*.cpp' -o -name '*.h' -o -name '*.py' -o -name 'CMakeLists.txt' \) -print0 | sort -z | xargs -0 sha1sum | sha1sum | cut -d' ' -f1)
if [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$WANT" ] && ls "$OUT"/TensorFrost/TensorFrost*.so >/dev/null 2>&1; then
  echo "[oracle] up to date"; exit 0
fi
rm -rf "$SCRATCH"; mkdir -p "$SCRATCH"
cp -r "$REF" "$SCRATCH/src"; chmod -R u+w "$SCRATCH/src"
S=$SCRATCH/src
# glad's generator otherwise downloads gl.xml; REPRODUCIBLE uses the vendored spec.
sed -i 's/glad_gl_core_46 SHARED API/glad_gl_core_46 STATIC REPRODUCIBLE API/' "$S/TensorFrost/CMakeLists.txt"
cmake -S "$S" -B "$SCRATCH/build" -G Ninja -DCMAKE_BUILD_TYPE=Release \
  -DGLFW_BUILD_X11=OFF -DGLFW_BUILD_WAYLAND=OFF -DCMAKE_POSITION_INDEPENDENT_CODE=ON \
  --compile-no-warning-as-error \
  -DCMAKE_CXX_FLAGS="-DGLFW_INCLUDE_NONE -fkeep-inline-functions" > "$SCRATCH/cmake.log" 2>&1
ninja -C "$SCRATCH/build" TensorFrost > "$SCRATCH/ninja.log" 2>&1
rm -rf "$OUT/TensorFrost"; mkdir -p "$OUT/TensorFrost"
cp "$S"/Python/TensorFrost/*.py "$OUT/TensorFrost/"
cp "$S"/Python/TensorFrost/*.so* "$OUT/TensorFrost/"
echo "$WANT" > "$STAMP"
echo "[oracle] built $(ls "$OUT"/TensorFrost/*.so)"
