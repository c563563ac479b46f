#!/usr/bin/env bash
# A/B of the split extrema / gradient pipeline against the fused round-1 kernel (development aid).
set -u
mkdir -p gpurun_out
for peak in 0 2; do
  echo "== split, peak $peak"; timeout 300 python tools/quick_bench.py 1920 1080 64 $peak 2>&1 | tail -3
  echo "== fused, peak $peak"; NM_EXTREMA_FUSED=1 timeout 300 python tools/quick_bench.py 1920 1080 64 $peak 2>&1 | tail -3
done
