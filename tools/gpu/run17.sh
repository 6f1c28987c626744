mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "synthetic or golden or many_candidate" 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --no-det > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
tail -1 gpurun_out/r02l_bench.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02l_bench.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
PY
