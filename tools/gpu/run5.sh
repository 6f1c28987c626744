mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "synthetic or many_candidate or golden or corner or streamed or pipelined" 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -3 gpurun_out/f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
print('other', {k:(v['ms_per_step'], v['stage_ms']) for k,v in d['other_modes'].items()})
PY
