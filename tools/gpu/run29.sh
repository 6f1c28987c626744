mkdir -p gpurun_out
nvidia-smi -L | head -3
( time timeout 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_shapes.py -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tail -7
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-det > gpurun_out/r02x_bench_n2.json 2> gpurun_out/r02x_bench_n2.err
echo "rc=$?"; grep -v "^\*\|OMP\|Percent" gpurun_out/r02x_bench_n2.err | tail -6
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02x_bench_n2.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
print('e2e', {k:v for k,v in (d['e2e'] or {}).items() if k not in ('pack_note',)})
print('parity', d['parity'])
print('extra', {k:(round(v['ms_per_step'],3), v['stage_ms']) for k,v in d['extra'].items()})
PY
