mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_properties.py -x -q 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -2 gpurun_out/r02e_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02e_bench.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
print('other', {k:(v['ms_per_step'], v['stage_ms']) for k,v in d['other_modes'].items()})
print('roofline', {k: d['roofline'][k] for k in ('kernel','frac','achieved','algorithmic_bytes','kernel_ms')}, d['roofline']['whole_pass_survey_formula'])
PY
