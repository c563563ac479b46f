#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own CUDA sources, from where they lie under
# /root/reference (never copied into this repo), plus oracle/ref_driver.cu, into
# oracle/_ref/libnmref.so for sm_100a.  The reference's CMake build cannot be used
# (FindCUDA-era, -arch=sm_30, CUDA-samples headers), so the sources
# are compiled directly with nvcc and two shim headers (oracle/shim/).
#
# One source-level fix is applied on the fly, to a temporary copy that is deleted
# again: src/gpu/kernels/match.cu:7-11 defines CHUNK=16 in the device pass and
# CHUNK=4 in the host pass on sm>=50, so the launch configuration disagrees with the
# kernel (out-of-bounds shared memory).  CHUNK is forced to 4 in both passes; the
# computed distances do not depend on CHUNK.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${NM_REFERENCE_DIR:-/root/reference}/src"
OUT="$HERE/_ref"
if [ ! -d "$REF/gpu/kernels" ]; then
    echo "[build_ref] $REF not present: keeping any prebuilt $OUT/libnmref.so" >&2
    exit 0
fi
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT"
ARCH="-gencode arch=compute_100a,code=sm_100a"
INC="-I$HERE/shim -I$REF/utils -I$REF/gpu/kernels -I$REF/gpu/utils -I$REF/gpu/sift"
FLAGS="$ARCH -std=c++17 -O2 -DUSE_CUDA -Xcompiler -fPIC -w $INC"
sed -e 's/#define CHUNK 16/#define CHUNK 4/' "$REF/gpu/kernels/match.cu" > "$TMP/match.cu"
# orientation.cu is included textually by ref_driver.cu (see there)
SRCS=(convolution downsample cudamath keypoint descriptor transpose bgra_2_gray cast undistort resample)
pids=()
for s in "${SRCS[@]}"; do
    $NVCC $FLAGS -c "$REF/gpu/kernels/$s.cu" -o "$TMP/$s.o" & pids+=($!)
done
$NVCC $FLAGS -c "$TMP/match.cu" -o "$TMP/match.o" & pids+=($!)
for s in pyramidata siftdata siftfunctions; do
    $NVCC $FLAGS -c "$REF/gpu/sift/$s.cu" -o "$TMP/$s.o" & pids+=($!)
done
$NVCC $FLAGS -c "$REF/gpu/utils/cudatex2D.cu" -o "$TMP/cudatex2D.o" & pids+=($!)
$NVCC $FLAGS -c "$HERE/ref_driver.cu" -o "$TMP/ref_driver.o" & pids+=($!)
$NVCC $FLAGS -c "$HERE/ref_preprocess_driver.cu" -o "$TMP/ref_preprocess_driver.o" & pids+=($!)
$NVCC $FLAGS -c "$HERE/ref_mosaic_driver.cu" -o "$TMP/ref_mosaic_driver.o" & pids+=($!)
# ransac.cu (+ svd.cu) is included textually by ref_ransac_driver.cu (see there)
$NVCC $FLAGS -c "$HERE/ref_ransac_driver.cu" -o "$TMP/ref_ransac_driver.o" & pids+=($!)
for p in "${pids[@]}"; do wait "$p"; done
$NVCC $ARCH -shared -o "$OUT/libnmref.so" "$TMP"/*.o
echo "[build_ref] built $OUT/libnmref.so"
