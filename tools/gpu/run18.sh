mkdir -p gpurun_out
for nb in 1 2 3 0; do
SFB200_SCATTER1_BLOCKS=$nb timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-e2e --no-cpu-baseline --no-det 2>/dev/null | python -c "
import json,sys
d=[json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
print('blocks $nb ms/step', round(d['ms_per_step'],3), d['stage_ms'])"
done
