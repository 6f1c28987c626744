mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/d_tests.log
tail -5 gpurun_out/d_tests.log
python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err
tail -3 gpurun_out/d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='pack_note'})
print('parity', d['parity']['status'], d['parity'].get('global_kernels'))
print('other', d['other_modes'])
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('per_kernel',)})
PY
