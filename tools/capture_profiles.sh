#!/usr/bin/env bash
# Round-2 profile capture on the GPU box (gpurun): the plain bench line, the ncu launch list of the same command, and
# one `--set full` capture per hot kernel (octave-0 launch of the second iteration of a 64-frame 1080p run), exported to
# CSV pages on the box (gpurun brings back at most 64 MiB, the .ncu-rep files are deleted).
set -u
O=gpurun_out
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --no-match"
$BENCH > $O/r02_bench_sift_only.json 2> $O/r02_bench_sift_only.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/r02_launches_bench.csv $BENCH > $O/r02_ncu_launches.log 2>&1
QB="python tools/quick_bench.py 1920 1080 64"
$QB > $O/r02_qb64.log 2>&1 || exit 1
cap() {   # name, regex on the MANGLED kernel name, launches to skip
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"$2" -s $3 -c 1 -o $O/r02_$1 $QB > $O/r02_ncu_$1.log 2>&1
  ncu -i $O/r02_$1.ncu-rep --page raw --csv > $O/r02_$1_raw.csv 2>/dev/null
  ncu -i $O/r02_$1.ncu-rep --page source --csv > $O/r02_$1_source.csv 2>/dev/null
  rm -f $O/r02_$1.ncu-rep
  python tools/ncu_source_top.py $O/r02_$1_source.csv 40 > $O/r02_$1_source_top.txt 2>&1
  rm -f $O/r02_$1_source.csv
}
cap blur_stream5 blur_stream_kernelILi5E 2
cap blur_stream7 blur_stream_kernelILi7E 4
cap blur_stream8 blur_stream_kernelILi8E 2
cap blur_stream10 blur_stream_kernelILi10E 2
cap blur_stream13 blur_stream_kernelILi13E 2
cap extrema 14extrema_kernel 6
cap refine_list refine_list_kernel 6
cap gradmap gradmap_kernel 6
cap orient 13orient_kernel 1
cap describe describe_fast_kernel 1
cap kprefine kprefine_kernel 1
M="python tools/match_once.py 100000 100000 1"
$M > $O/r02_match_once.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_scan_kernel" -s 3 -c 1 -o $O/r02_tc_scan $M > $O/r02_ncu_tc_scan.log 2>&1
ncu -i $O/r02_tc_scan.ncu-rep --page raw --csv > $O/r02_tc_scan_raw.csv 2>/dev/null
ncu -i $O/r02_tc_scan.ncu-rep --page source --csv > $O/r02_tc_scan_source.csv 2>/dev/null
rm -f $O/r02_tc_scan.ncu-rep
python tools/ncu_source_top.py $O/r02_tc_scan_source.csv 40 > $O/r02_tc_scan_source_top.txt 2>&1
rm -f $O/r02_tc_scan_source.csv
ls -la $O | tail -40
du -sh $O
