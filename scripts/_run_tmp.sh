N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r01z_bench_n$N.json 2> gpurun_out/r01z_bench_n$N.err; echo "n$N exit $?"
tail -2 gpurun_out/r01z_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r01z_bench_n$N.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['ms_per_step_eager_instrumented'], d['e2e'], d['matches_per_step'])
PY
