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
# orientation.cu is included textually by ref_driver.cu (see there).  A second, patched temporary copy gives the
# reference's PUBLIC orientation kernel in a form that terminates on sm_70+: kernel_orientations_optim calls
# __syncthreads() inside `if (index < NBINS - 1)` (orientation.cu:68-86), which deadlocks under independent thread
# scheduling.  The copy hoists the two barriers out of the branch and keeps every arithmetic statement (the
# 10-pixel window clamp :29-30, the in-place update of hist[35] by thread 0 :78-80 with its race against thread 34,
# first-two-peaks :118-127); the symbols get a _hoisted suffix so that both variants link into one library.
ORI="$REF/gpu/kernels/orientation.cu"
l68="$(sed -n 68p "$ORI")"; l71="$(sed -n 71p "$ORI")"; l81="$(sed -n 81p "$ORI")"; l83="$(sed -n 83p "$ORI")"; l84="$(sed -n 84p "$ORI")"
case "$l68" in *"if (index < NBINS - 1) {"*) ;; *) echo "[build_ref] orientation.cu:68 is not what the patch expects" >&2; exit 1;; esac
case "$l71" in *"for (int iter = 0; iter < 6; ++iter) {"*) ;; *) echo "[build_ref] orientation.cu:71 unexpected" >&2; exit 1;; esac
case "$l81$l84" in *"__syncthreads();"*"__syncthreads();"*) ;; *) echo "[build_ref] orientation.cu:81/84 unexpected" >&2; exit 1;; esac
case "$l83" in *"hist[index] = temp[index];"*) ;; *) echo "[build_ref] orientation.cu:83 unexpected" >&2; exit 1;; esac
sed -e '68s/.*/    { const bool nm_act = (index < NBINS - 1);/' \
    -e '71s/.*/        for (int iter = 0; iter < 6; ++iter) { if (nm_act) {/' \
    -e '81s/.*/            } __syncthreads();/' \
    -e '83s/.*/            if (nm_act) hist[index] = temp[index];/' \
    -e 's/kernel_orientations_optim/kernel_orientations_optim_hoisted/g' \
    -e 's/kernel_orientations_naive/kernel_orientations_naive_unused/g' \
    -e 's/detect_orientations/detect_orientations_hoisted/g' "$ORI" > "$TMP/orientation_hoisted.cu"
SRCS=(convolution downsample cudamath keypoint descriptor transpose bgra_2_gray cast undistort resample)
pids=()
for s in "${SRCS[@]}"; do
    $NVCC $FLAGS -c "$REF/gpu/kernels/$s.cu" -o "$TMP/$s.o" & pids+=($!)
done
$NVCC $FLAGS -c "$TMP/match.cu" -o "$TMP/match.o" & pids+=($!)
$NVCC $FLAGS -c "$TMP/orientation_hoisted.cu" -o "$TMP/orientation_hoisted.o" & pids+=($!)
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
