mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r03f_bench_c3_full.json 2> gpurun_out/r03f_bench.err ) 2> gpurun_out/r03f_time.txt
tail -3 gpurun_out/r03f_time.txt; grep -v "^\*\|OMP\|Percent" gpurun_out/r03f_bench.err | tail -5
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r03f_bench_c3_full.json') if l.startswith('{')][-1]
print('ms/step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'])
print('stage', d['stage_ms'])
print('parity', d['parity'])
print('e2e', {k:v for k,v in d['e2e'].items() if k!='pack_note'})
print('roofline', {k:v for k,v in d['roofline'].items() if k not in ('per_kernel','whole_pass_survey_formula_r01_work','whole_pass_streamed_bytes')})
print('per_kernel', {k:(round(x['ms'],3), round(x['frac'],3)) for k,x in d['roofline']['per_kernel'].items()})
print('cpu', {k:v for k,v in d['cpu_baseline'].items() if k!='sample'})
print('decode job', d['decode']['job']['seconds'], d['decode']['native_lines_per_s'])
print('other', {k:(round(v['ms_per_step'],3)) for k,v in d['other_modes'].items()})
print('extra', {k:(round(v['ms_per_step'],3), v['stage_ms']) for k,v in d['extra'].items()})
print('clocks', d['clocks'])
PY
