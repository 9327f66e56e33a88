TAG=${1:-b}
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('${TAG}: count %.4e (%.2f ms)  ids %.4e (%.2f ms)'%(d['value'], d['ms_per_step'], d['ids_mode']['value'], d['ids_mode']['ms_per_step']))
PY
