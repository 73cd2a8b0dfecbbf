timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r01y_bench.json 2> gpurun_out/r01y_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r01y_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r01y_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), d['e2e'])
PY
