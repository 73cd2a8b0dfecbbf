timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r01u_bench.json 2>gpurun_out/r01u_bench.err; tail -3 gpurun_out/r01u_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01u_bench.json'))
print(d['value'], d['ms_per_step'], d['execution'], d['value_eager_instrumented'], d['next_rows'])
PY
