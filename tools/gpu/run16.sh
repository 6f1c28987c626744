mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_tile_share" -s 1 -c 1 -o gpurun_out/r02k_share python bench.py --scale 0.4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02k_ncu.log 2>&1
tail -2 gpurun_out/r02k_ncu.log | cut -c1-300
