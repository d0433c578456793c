#!/bin/bash
# ncu `--set full` tables of the kernels changed late in round 2 (events kernel, --dtw-std pair kernel, pair kernel with
# 13 rows per lane), on small workloads.  Run under gpurun, one GPU.  Output: gpurun_out/<tag>_*.  Usage: tools/ncu_capture_small.sh r02b
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
set -x
timeout 300 python tools/perf_sweep.py c2 --reads 4096 > $out/${tag}_small_plain.txt 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none -k regex:'sf_events_kernel|sf_dtw_pair_kernel|sf_trace_pair_kernel' -c 3 \
    -o $out/${tag}_prof_c2 -f python tools/perf_sweep.py c2 --reads 4096 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'sf_dtw_pair_kernel' -c 1 \
    -o $out/${tag}_prof_std -f python tools/perf_sweep.py c5 --only-std >> $out/${tag}_small_plain.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'sf_dtw_pair_kernel' -c 1 \
    -o $out/${tag}_prof_q200 -f python tools/perf_sweep.py c4 --reads 2048 -q 200 >> $out/${tag}_small_plain.txt 2>&1
python tools/ncu_summary.py $out/${tag}_prof_c2.ncu-rep $out/${tag}_prof_std.ncu-rep $out/${tag}_prof_q200.ncu-rep > $out/${tag}_ncu_summary_tables.md 2> $out/${tag}_ncu_summary_err.txt
rm -f $out/*.ncu-rep
ls -la $out | grep ${tag}_
