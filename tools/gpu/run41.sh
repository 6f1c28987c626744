mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --scale 0.2 --steps 3 --warmup 3 --no-det --no-extras > gpurun_out/r03g_n2_small.json 2> gpurun_out/r03g_n2_small.err
echo "rc=$?"; grep -v "^\*\|OMP\|Percent" gpurun_out/r03g_n2_small.err | tail -4
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r03g_n2_small.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'bound', d.get('host_cores_bound'), 'parity', d['parity']['status'])
print('stage', d['stage_ms'])
print('e2e', {k:v for k,v in (d['e2e'] or {}).items() if k not in ('pack_note',)})
PY
