#!/usr/bin/env bash
# Round-2 profile capture on the GPU box (gpurun): the plain bench line, the ncu launch list of the same command, and
# one `--set full` capture per hot kernel.  Outputs go to gpurun_out/r02_*; summaries are made here afterwards.
set -u
O=gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --no-match"
$BENCH > $O/r02_bench_sift_only.json 2> $O/r02_bench_sift_only.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/r02_launches_bench.csv $BENCH > $O/r02_ncu_launches.log 2>&1
QB="python tools/quick_bench.py 1920 1080 16"
$QB > $O/r02_qb16.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"blur_strip_kernel|extrema_kernel|refine_list|gradmap_kernel|orient_kernel|describe_fast|kprefine|emit_kernel|rank_kernel" -s 57 -c 57 -o $O/r02_prof_sift $QB > $O/r02_ncu_sift.log 2>&1
M="python tools/match_once.py 100000 100000 1"
$M > $O/r02_match_once.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_scan_kernel|tc_rerank_kernel" -s 2 -c 3 -o $O/r02_prof_match $M > $O/r02_ncu_match.log 2>&1
tail -2 $O/r02_ncu_sift.log $O/r02_ncu_match.log
