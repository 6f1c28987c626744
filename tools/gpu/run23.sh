mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic or golden or transfer or reproducible or degenerate or streamed" 2>&1 | tail -4 ) 2>&1 | tail -8
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-det --no-extras"
for v in default o1 o2; do
  if [ $v = default ]; then unset SFB200_LIB; E2E=""; else export SFB200_LIB=$PWD/build/var/lib_$v.so; E2E="--no-e2e"; fi
  timeout 400 python bench.py $Q $E2E > gpurun_out/r02r_quick_$v.json 2> gpurun_out/r02r_quick_$v.err || { echo "bench $v failed"; tail -5 gpurun_out/r02r_quick_$v.err; }
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=[json.loads(l) for l in open(f'gpurun_out/r02r_quick_{v}.json') if l.startswith('{')][-1]
    print(v, 'ms/step', round(d['ms_per_step'],3), d['stage_ms'])
    if d.get('e2e'): print('e2e', {k:x for k,x in d['e2e'].items() if k!='pack_note'})
except Exception as e: print(v, 'no line', e)
PY
done
