mkdir -p gpurun_out
for mode in "--steps 20 --warmup 3 --no-cpu-baseline --no-det --no-extras --no-e2e" "--steps 10 --warmup 3 --no-det --no-extras --no-e2e"; do
  timeout 400 python bench.py $mode > gpurun_out/r03e_diag.json 2> gpurun_out/r03e_diag.err || { echo "bench failed"; tail -5 gpurun_out/r03e_diag.err; }
  python - "$mode" <<'PY'
import json,sys
d=[json.loads(l) for l in open('gpurun_out/r03e_diag.json') if l.startswith('{')][-1]
print(sys.argv[1], '->', round(d['ms_per_step'],3), d['stage_ms'])
PY
done
