#!/usr/bin/env bash
# Re-capture after a change to one kernel: bench line, launch list, and the named `--set full` captures.
# usage: tools/capture_final.sh "<name> <mangled regex> <skip>" ...
set -u
O=gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --no-match"
$BENCH > $O/r02_bench_sift_only.json 2> $O/r02_bench_sift_only.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/r02_launches_bench.csv $BENCH > $O/r02_ncu_launches.log 2>&1
QB="python tools/quick_bench.py 1920 1080 64"
for spec in "$@"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"$2" -s $3 -c 1 -o $O/r02_$1 $QB > $O/r02_ncu_$1.log 2>&1
  ncu -i $O/r02_$1.ncu-rep --page raw --csv > $O/r02_$1_raw.csv 2>/dev/null
  ncu -i $O/r02_$1.ncu-rep --page source --csv > $O/r02_$1_source.csv 2>/dev/null
  rm -f $O/r02_$1.ncu-rep
  python tools/ncu_source_top.py $O/r02_$1_source.csv 40 > $O/r02_$1_source_top.txt 2>&1
  rm -f $O/r02_$1_source.csv
done
