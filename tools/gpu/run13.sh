mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_properties.py -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --no-det > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
tail -2 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02f_bench.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
PY
ncu --set full --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum --clock-control none --import-source on -k regex:"k_match_steps|k_classify|k_pass1_scatter|k_pass2_resolve|k_pass2_scatter" -s 25 -c 7 -o gpurun_out/r02f_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02f_ncu.log 2>&1
tail -1 gpurun_out/r02f_ncu.log | cut -c1-200
