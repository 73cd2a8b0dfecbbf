#!/bin/bash
# final 8-GPU lines of round 2 (one box):  gpurun --gpus 8 --timeout 600 -- 'bash scripts/gpu_r2_n8_final.sh'
OUT=gpurun_out; mkdir -p $OUT
Q="--no-cpu-baseline --no-gpu-baselines --no-next-rows --no-sweep"
run() {  # name nproc args...
  local name=$1 n=$2; shift 2
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n $Q "$@" \
      > $OUT/r2f_${name}.json 2> $OUT/r2f_${name}.err
  python scripts/bench_digest.py $OUT/r2f_${name}.json 2>/dev/null | head -3
}
run bench_n8 8 --steps 20 --warmup 5
run bench_n2 2 --steps 20 --warmup 5 --no-e2e
run indoor_p4_n8 8 --config indoor --size 640x480 --pairs 4 --steps 10 --warmup 3 --no-e2e
run s832_g16_n8 8 --size 832 --global-pairs 16 --steps 10 --warmup 3 --no-e2e
