timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r01w_bench.json 2> gpurun_out/r01w_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r01w_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'eager', round(d['ms_per_step_eager_instrumented'],4), d['execution'], d['e2e'], d['next_rows'].get('quadtree_attention_tokens'))
PY
