mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transfer or streamed or shards" 2>&1 | tail -2 )
python tools/expand_bench.py 0.25 2>&1 | tail -1
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-det --no-extras"
timeout 400 python bench.py $Q > gpurun_out/r03c_e2e.json 2> gpurun_out/r03c_e2e.err || { echo "bench failed"; tail -5 gpurun_out/r03c_e2e.err; }
python - <<'PY'
import json,sys
d=[json.loads(l) for l in open('gpurun_out/r03c_e2e.json') if l.startswith('{')][-1]
print('ms/step', round(d['ms_per_step'],3), 'e2e', {k:x for k,x in d['e2e'].items() if k!='pack_note'})
PY
