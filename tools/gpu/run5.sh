mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "synthetic or many_candidate or golden or corner or streamed or pipelined" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -2 gpurun_out/f_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/f_bench.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
print('other', {k:(v['ms_per_step'], v['stage_ms']) for k,v in d['other_modes'].items()})
PY
SFB200_TILE_KERNELS=1 python bench.py --config c4 --steps 3 --warmup 2 --no-extras --no-e2e --no-cpu-baseline --no-det 2> gpurun_out/h.err | grep "^{" > gpurun_out/h_c4_tile.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/h_c4_tile.json').read())
print('c4 tile kernels ms/step', d['ms_per_step'], d['stage_ms'])
PY
