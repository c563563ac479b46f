#!/usr/bin/env bash
# A/B of the gradient-map kernel variants (development aid)
for v in 1 7; do
  echo "== NM_GRADMAP=$v"; NM_GRADMAP=$v timeout 300 python tools/quick_bench.py 1920 1080 64 0 2>&1 | tail -2 | head -1
done
