mkdir -p gpurun_out
free -g | head -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras > gpurun_out/g_bench_n2.json 2> gpurun_out/g_bench_n2.err
echo "rc=$?"; grep -v "^\*\|OMP" gpurun_out/g_bench_n2.err | tail -12
python - <<'PY'
import json
d=json.load(open('gpurun_out/g_bench_n2.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'])
print('stage', d['stage_ms'])
print('e2e', {k:v for k,v in (d['e2e'] or {}).items() if k not in ('pack_note',)})
print('parity', d['parity']['status'])
print('other', {k:(v['ms_per_step']) for k,v in d['other_modes'].items()})
PY
