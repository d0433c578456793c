#!/bin/bash
# Captures the ncu evidence of a round on the GPU box (run under gpurun, one GPU): launch lists with device times and
# `--set full` reports of the hot kernels on the C4 / C2 / C5 shapes.  Output: gpurun_out/<tag>_*.  Usage: tools/ncu_capture.sh r02
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
D=/tmp/sigfish_b200/cli_e2e
CLI=sigfish_b200/sigfish-b200
set -x
# files of the C4 shape (one batch of 8288 reads) and a plain run
timeout 300 python tools/cli_e2e.py --reads 8288 --gpus 1 > $out/${tag}_ncu_plain.txt 2>&1 || exit 1
ARGS="dtw $D/ref.fa $D/reads.blow5 --kmer-model $D/model.txt -K 8288 -B 100G -t 16 --gpus 1 -o $D/ncu.paf"
timeout 120 $CLI $ARGS > /dev/null 2>&1 || exit 1
# 1. every launch of the command line with its device time
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches_cli_c4.csv $CLI $ARGS > /dev/null 2>&1
# 2. full reports of the kernels of that run
timeout 900 ncu --set full --clock-control none \
    -k regex:'sf_dtw_pair_kernel|sf_inflate_kernel|sf_signal_kernel|sf_events_kernel|sf_trace_pair_kernel|sf_verify_kernel|sf_ref_stats_kernel' \
    -o $out/${tag}_prof_cli_c4 -f $CLI $ARGS > /dev/null 2>&1
# 3. --sam: the path kernel
timeout 120 $CLI $ARGS --sam > /dev/null 2>&1 && \
timeout 600 ncu --set full --clock-control none -k regex:sf_path_kernel -c 1 -o $out/${tag}_prof_path -f $CLI $ARGS --sam > /dev/null 2>&1
# 4. C2 and C5 shapes (tools/perf_sweep.py: submit, 3 resubmits, submit per shape)
timeout 300 python tools/perf_sweep.py c2 c5 --reads 4096 > $out/${tag}_ncu_sweep_plain.txt 2>&1 && \
timeout 900 ncu --set full --clock-control none -k regex:'sf_dtw_pair_kernel|sf_dtw_score_kernel|sf_events_kernel|sf_trace' -c 14 \
    -o $out/${tag}_prof_c2_c5 -f python tools/perf_sweep.py c2 c5 --reads 4096 > /dev/null 2>&1
# the reports are large (tens of MB each; gpurun brings back 64 MiB at most): keep the metric tables, drop the reports
python tools/ncu_summary.py $out/${tag}_prof_cli_c4.ncu-rep $out/${tag}_prof_path.ncu-rep $out/${tag}_prof_c2_c5.ncu-rep > $out/${tag}_ncu_summary.md 2> $out/${tag}_ncu_summary_err.txt
for r in prof_cli_c4 prof_path prof_c2_c5; do
    ncu -i $out/${tag}_$r.ncu-rep --page raw --csv > $out/${tag}_$r.raw.csv 2>/dev/null
done
rm -f $out/*.ncu-rep
ls -la $out | grep ${tag}_
