mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic or golden or reproducible or degenerate or explicit or many_candidate" 2>&1 | tail -4 ) 2>&1 | tail -8
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-det --no-extras --no-e2e"
for v in default f0; do
  if [ $v = default ]; then unset SFB200_LIB; else export SFB200_LIB=$PWD/build/var/lib_$v.so; fi
  timeout 400 python bench.py $Q > gpurun_out/r02z_quick_$v.json 2> gpurun_out/r02z_quick_$v.err || { echo "bench $v failed"; tail -5 gpurun_out/r02z_quick_$v.err; }
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=[json.loads(l) for l in open(f'gpurun_out/r02z_quick_{v}.json') if l.startswith('{')][-1]
    print(v, 'ms/step', round(d['ms_per_step'],3), d['stage_ms'])
except Exception as e: print(v, 'no line', e)
PY
done
unset SFB200_LIB
timeout 300 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --no-det --no-extras --no-e2e > gpurun_out/r02z_c4.json 2> gpurun_out/r02z_c4.err
python - <<'PY'
import json
d=[json.loads(l) for l in open('gpurun_out/r02z_c4.json') if l.startswith('{')][-1]
print('c4 ms/step', round(d['ms_per_step'],3), d['stage_ms'])
PY
