mkdir -p gpurun_out
SFB200_TILE_KERNELS=1 ncu --set full --clock-control none --import-source on -k regex:k_tile_pass -s 2 -c 2 -o gpurun_out/r02d_tile python bench.py --scale 0.4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02d_ncu.log 2>&1
tail -2 gpurun_out/r02d_ncu.log | cut -c1-300
