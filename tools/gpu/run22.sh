mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -m gpu -x -q -k "synthetic or golden or many_candidate or degenerate or shape or reproducible" 2>&1 | tail -4 ) 2>&1 | tail -8
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras"
for v in default w1 w2; do
  if [ $v = default ]; then unset SFB200_LIB; else export SFB200_LIB=$PWD/build/var/lib_$v.so; fi
  timeout 300 python bench.py $Q > gpurun_out/r02q_quick_$v.json 2> gpurun_out/r02q_quick_$v.err || { echo "bench $v failed"; tail -5 gpurun_out/r02q_quick_$v.err; }
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=[json.loads(l) for l in open(f'gpurun_out/r02q_quick_{v}.json') if l.startswith('{')][-1]
    print(v, 'ms/step', round(d['ms_per_step'],3), d['stage_ms'])
except Exception as e: print(v, 'no line', e)
PY
done
unset SFB200_LIB
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_tile_share|k_classify" -s 8 -c 3 -o gpurun_out/r02q_full python bench.py --scale 0.4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02q_ncu.log 2>&1
tail -2 gpurun_out/r02q_ncu.log | cut -c1-200
