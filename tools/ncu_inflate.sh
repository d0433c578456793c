mkdir -p gpurun_out
timeout 100 python tools/decode_bench.py --reads 16384 > gpurun_out/ncu_inflate_plain.txt 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sf_inflate_kernel -s 1 -c 1 -o gpurun_out/inflate -f python tools/decode_bench.py --reads 16384 > /dev/null 2>&1
ncu -i gpurun_out/inflate.ncu-rep --page source --csv > gpurun_out/r02_inflate_source.csv 2>/dev/null
ncu -i gpurun_out/inflate.ncu-rep --page raw --csv > gpurun_out/r02_inflate_raw.csv 2>/dev/null
rm -f gpurun_out/inflate.ncu-rep
ls -la gpurun_out | grep inflate
