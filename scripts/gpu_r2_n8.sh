#!/bin/bash
# 8-GPU pass of round 2 (one box):  gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_r2_n8.sh TAG'
#   cfg 2 weak scaling at N = 8 (with the host-fed e2e leg), cfg 4 (indoor 640x480, 32 pairs over 8 GPUs), cfg 5 (16 pairs, strong scaling)
TAG=${1:-r2n8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt; nvidia-smi topo -m >> $OUT/${TAG}_gpus.txt 2>&1
Q="--no-cpu-baseline --no-gpu-baselines --no-next-rows --no-sweep"
run() {  # name nproc args...
  local name=$1 n=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n $Q "$@" \
      > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python scripts/bench_digest.py $OUT/${TAG}_${name}.json 2>&1 | head -4
}
run bench_n8 8 --steps 20 --warmup 5
run indoor_p4_n8 8 --config indoor --size 640x480 --pairs 4 --steps 10 --warmup 3 --no-e2e
for S in 832 1152; do
  for N in 2 4 8; do run s${S}_g16_n$N $N --size $S --global-pairs 16 --steps 10 --warmup 3 --no-e2e; done
done
for S in 512 1024; do run s${S}_g16_n8 8 --size $S --global-pairs 16 --steps 10 --warmup 3 --no-e2e; done
ls -la $OUT | tail -20
