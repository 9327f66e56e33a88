# usage: bash tools/run_quick.sh TAG [pytest-args]  — GPU parity tests, then one short bench (no e2e / cpu legs)
TAG=${1:-q}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('count %.4e (%.2f ms)  ids %.4e (%.2f ms)'%(d['value'], d['ms_per_step'], d['ids_mode']['value'], d['ids_mode']['ms_per_step']))
PY
