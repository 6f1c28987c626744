mkdir -p gpurun_out
SFB200_TILE_KERNELS=1 python bench.py --config c4 --steps 3 --warmup 2 --no-extras --no-e2e --no-cpu-baseline --no-det 2> gpurun_out/h.err | grep "^{" > gpurun_out/h_c4_tile.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/h_c4_tile.json').read())
print('c4 tile kernels ms/step', d['ms_per_step'], d['stage_ms'])
PY
