mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r02m_bench_c3_full.json 2> gpurun_out/r02m_bench.err ) 2> gpurun_out/r02m_time.txt
tail -3 gpurun_out/r02m_time.txt; grep -v "^\*\|OMP" gpurun_out/r02m_bench.err | tail -5
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02m_bench_c3_full.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'])
print('stage', d['stage_ms'])
print('parity', d['parity'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='pack_note'})
print('roofline', {k:v for k,v in d['roofline'].items() if k!='per_kernel'})
print('per_kernel', d['roofline']['per_kernel'])
print('cpu', d['cpu_baseline'])
print('other', {k:(round(v['ms_per_step'],3)) for k,v in d['other_modes'].items()})
print('extra', {k:(round(v['ms_per_step'],3), v['stage_ms']) for k,v in d['extra'].items()})
PY
