mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "synthetic or many_candidate or match_lists or corner or step_index or golden or multiword or regrows or degenerate or explicit or streamed or pipelined or reproducible" 2>&1 | tail -5
timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --no-det > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
tail -2 gpurun_out/r02h_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02h_bench.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
PY
timeout 200 python bench.py --config c4 --steps 3 --warmup 2 --no-extras --no-e2e --no-cpu-baseline --no-det 2> gpurun_out/r02h_c4.err | grep "^{" > gpurun_out/r02h_c4.json
tail -2 gpurun_out/r02h_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02h_c4.json').read())
print('c4 ms/step', d['ms_per_step'], d['stage_ms'])
PY
