mkdir -p gpurun_out
# 1. the whole GPU suite, with the slowest tests listed
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -22 ) 2>&1 | tail -28
# 2. smoke
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
# 3. ncu full capture of the four big kernels of one step at full scale (after the plain run of the same command)
Q="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras"
timeout 300 python bench.py $Q > gpurun_out/r02w_plain.json 2> gpurun_out/r02w_plain.err && \
timeout 500 ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum --clock-control none --import-source on -k regex:"k_tile_share|k_match_steps|k_classify|k_tile_scatter1" -s 15 -c 5 -o gpurun_out/r02w_full python bench.py $Q > gpurun_out/r02w_ncu.log 2>&1
tail -2 gpurun_out/r02w_ncu.log | cut -c1-160
ls -la gpurun_out/r02w*
