timeout 900 python -m pytest tests/test_gpu_vs_reference_ext.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r01y_bench.json 2> gpurun_out/r01y_bench.err; echo "bench exit $?"
tail -3 gpurun_out/r01y_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r01y_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value'],1), 'eager', round(d['value_eager_instrumented'],1), d['e2e'])
print(d['gpu_torch_baseline']); print(d['gpu_reference_kernels_baseline']); print(d['cpu_baseline'])
PY
