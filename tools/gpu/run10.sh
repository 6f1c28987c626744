mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02c_tests.log; tail -4 gpurun_out/r02c_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum --clock-control none --import-source on -k regex:"k_match" -s 3 -c 1 -o gpurun_out/r02c_match python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02c_ncu.log 2>&1
tail -1 gpurun_out/r02c_ncu.log | cut -c1-100
