mkdir -p gpurun_out
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-det --no-extras"
for sh in 1 2; do
  timeout 400 python bench.py $Q --e2e-shards $sh > gpurun_out/r03a_e2e_s$sh.json 2> gpurun_out/r03a_e2e_s$sh.err || { echo "bench $sh failed"; tail -5 gpurun_out/r03a_e2e_s$sh.err; }
  python - $sh <<'PY'
import json,sys
v=sys.argv[1]
d=[json.loads(l) for l in open(f'gpurun_out/r03a_e2e_s{v}.json') if l.startswith('{')][-1]
print('shards', v, 'ms/step', round(d['ms_per_step'],3), 'e2e', {k:x for k,x in d['e2e'].items() if k!='pack_note'})
PY
done
