timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r01n_bench_n2.json 2>gpurun_out/r01n_bench_n2.err; tail -3 gpurun_out/r01n_bench_n2.err
python - <<'PY'
import json
txt=open('gpurun_out/r01n_bench_n2.json').read()
print(txt[:80])
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][0])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['execution'], d['value_eager_instrumented'], d['cuda_graph'], d['e2e'])
PY
