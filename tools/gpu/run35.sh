mkdir -p gpurun_out
python tools/expand_bench.py 0.25 2>&1 | tail -1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_expand_walks|k_scan_write" -s 6 -c 3 -o gpurun_out/r03b_expand python tools/expand_bench.py 0.25 > gpurun_out/r03b_ncu.log 2>&1
tail -1 gpurun_out/r03b_ncu.log | cut -c1-150
