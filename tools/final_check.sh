#!/bin/bash
# Last GPU call of round 2 (≈ 2 min of box time): the command line on BLOW5 inputs after the head decoder change, a short
# bench line, the ncu launch list of that bench command, and the 30 kb shape from files.  Output: gpurun_out/r02e_*.
out=gpurun_out
mkdir -p $out
K="dna_sp1_from_end or dna_short_reads_p200 or rna004_tx2000_invert or rna_sequin_default or rna004_tail16_auto or dna_synth48_from_end"
timeout 60 python -m pytest tests/test_gpu_cli.py -x -q -k "byte_identical and ($K)" > $out/r02e_pytest_cli_blow5.txt 2>&1
echo "pytest rc $?" >> $out/r02e_pytest_cli_blow5.txt
timeout 40 python -m pytest tests/test_gpu_blow5.py -x -q -k "reference_blow5_file or corrupt_record" >> $out/r02e_pytest_cli_blow5.txt 2>&1
echo "pytest rc $?" >> $out/r02e_pytest_cli_blow5.txt
BARGS="--shapes= --no-files --no-cpu-baseline --steps 3 --warmup 3"
timeout 60 python bench.py $BARGS > $out/r02e_bench_short.json 2> $out/r02e_bench_short.err || exit 1
timeout 40 python tools/cli_e2e.py --shape c2 --reads 1002848 --auto-batch > $out/r02e_cli_e2e_c2.json 2> $out/r02e_cli_e2e_c2.err
timeout 80 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02e_launches_bench.csv \
    python bench.py $BARGS > $out/r02e_bench_under_ncu.txt 2>&1
ls -la $out | grep r02e_
