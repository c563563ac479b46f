#!/usr/bin/env bash
# A/B of the descriptor kernels (development aid): NM_DESCRIBE = 0 (round 1), 16, 32
for v in 16 162; do
  echo "== NM_DESCRIBE=$v"; NM_DESCRIBE=$v timeout 300 python tools/quick_bench.py 1920 1080 64 0 2>&1 | tail -2
done
