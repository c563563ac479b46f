#!/usr/bin/env bash
# One `ncu --set full` capture of a kernel selected by a regex on its MANGLED name, exported to CSV pages on the box.
# usage: tools/cap_one.sh <tag> <mangled-name regex> <skip> [frames]
set -u
O=gpurun_out
QB="python tools/quick_bench.py 1920 1080 ${4:-64}"
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"$2" -s $3 -c 1 -o $O/$1 $QB > $O/ncu_$1.log 2>&1
ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
ncu -i $O/$1.ncu-rep --page source --csv > $O/$1_source.csv 2>/dev/null
rm -f $O/$1.ncu-rep
python tools/ncu_source_top.py $O/$1_source.csv 40 > $O/$1_source_top.txt 2>&1
tail -3 $O/ncu_$1.log
