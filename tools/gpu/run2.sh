mkdir -p gpurun_out
python -m pytest tests/test_gpu_properties.py tests/test_gpu_shapes.py -x -q 2>&1 | tail -30 > gpurun_out/b_tests.log
tail -5 gpurun_out/b_tests.log
python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
tail -3 gpurun_out/b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
print('e2e', d['e2e'])
print('parity', d['parity'])
print('other', d['other_modes'])
print('cpu', d['cpu_baseline'])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='per_kernel'})
PY
