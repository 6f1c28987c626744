mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "transfer or streamed or pipelined" 2>&1 | tail -2 )
nvidia-smi topo -m 2>/dev/null | head -8
python - <<'PY'
import os, sys
sys.path.insert(0, '.')
import torch
from strainflair_b200 import dist
print('cores before', len(os.sched_getaffinity(0)))
print('bound', dist.bind_to_gpu_numa(0))
PY
Q="--steps 10 --warmup 3 --no-cpu-baseline --no-det --no-extras"
for v in numa nonuma; do
  if [ $v = numa ]; then unset SFB200_NO_NUMA; else export SFB200_NO_NUMA=1; fi
  timeout 400 python bench.py $Q > gpurun_out/r02z_e2e_$v.json 2> gpurun_out/r02z_e2e_$v.err || { echo "bench $v failed"; tail -5 gpurun_out/r02z_e2e_$v.err; }
  python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=[json.loads(l) for l in open(f'gpurun_out/r02z_e2e_{v}.json') if l.startswith('{')][-1]
print(v, 'ms/step', round(d['ms_per_step'],3), 'bound', d.get('host_cores_bound'), 'e2e', {k:x for k,x in d['e2e'].items() if k!='pack_note'})
PY
done
python tools/h2d_bench.py 2>&1 | tail -3
