mkdir -p gpurun_out
nvidia-smi -L | head -2
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) 2>&1 | tail -9
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras"
for v in default v1 v2 v3; do
  if [ $v = default ]; then unset SFB200_LIB; else export SFB200_LIB=$PWD/build/var/lib_$v.so; fi
  timeout 300 python bench.py $Q > gpurun_out/r02p_quick_$v.json 2> gpurun_out/r02p_quick_$v.err || { echo "bench $v failed"; tail -5 gpurun_out/r02p_quick_$v.err; }
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=[json.loads(l) for l in open(f'gpurun_out/r02p_quick_{v}.json') if l.startswith('{')][-1]
    print(v, 'ms/step', round(d['ms_per_step'],3), d['stage_ms'])
    if v=='default': print(d['workload_stats']); print({k:(round(x['ms'],3), round(x['frac'],3)) for k,x in d['roofline']['per_kernel'].items()})
except Exception as e: print(v, 'no line', e)
PY
done
unset SFB200_LIB
# ncu: full set of the two kernels that changed, one launch each, at scale 0.4
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_tile_share|k_match_steps" -s 8 -c 2 -o gpurun_out/r02p_full python bench.py --scale 0.4 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-det --no-extras > gpurun_out/r02p_ncu.log 2>&1
tail -2 gpurun_out/r02p_ncu.log | cut -c1-200
ls -la gpurun_out | tail -8
