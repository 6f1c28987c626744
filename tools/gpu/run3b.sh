mkdir -p gpurun_out
SFB200_TILE_KERNELS=1 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 1 -c 1 -o gpurun_out/i_tile_c4 python bench.py --config c4 --scale 0.1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/i_ncu.log 2>&1
tail -2 gpurun_out/i_ncu.log | cut -c1-300
