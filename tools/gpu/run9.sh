mkdir -p gpurun_out
# launch list of one bench command (default engine), per-launch durations
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_c3_full.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02b_ncu1.log 2>&1
tail -1 gpurun_out/r02b_ncu1.log | cut -c1-200
# full capture of the kernels of one step, with the L2 atomic counters
ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,l1tex__m_l1tex2xbar_write_sectors_mem_global_op_red.sum,lts__t_sectors_srcunit_tex_op_red.sum --clock-control none --import-source on -k regex:"k_match|k_classify|k_pass1_scatter|k_pass2_resolve|k_pass2_scatter|k_path_reduce|k_collapse" -s 28 -c 8 -o gpurun_out/r02b_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02b_ncu2.log 2>&1
tail -1 gpurun_out/r02b_ncu2.log | cut -c1-200
ls -la gpurun_out | tail -5
