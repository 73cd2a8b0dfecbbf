#!/bin/bash
# Round-2 GPU-box pass.  Usage:  gpurun --timeout 1500 -- 'bash scripts/gpu_r2.sh TAG [stage ...]'
# stages: test (pytest -m gpu + smoke), bench (both arms), sweep (configs 2c / indoor / sizes), ncu (launch list + --set full captures)
# Everything lands in gpurun_out/ (kept under 64 MiB).
TAG=${1:-r2a}; shift
STAGES=${@:-test bench}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|NUMA" >> $OUT/${TAG}_gpu.txt
free -g | head -2 >> $OUT/${TAG}_gpu.txt
for st in $STAGES; do
case $st in
test)
  timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -15 $OUT/${TAG}_pytest.log
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
  ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
  tail -3 $OUT/${TAG}_bench.err
  python scripts/bench_digest.py $OUT/${TAG}_bench.json
  ;;
ref)
  timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
  cut -c1-300 $OUT/${TAG}_bench_ref.json
  ;;
simt)   # A/B: dense coarsest level on the fp32 SIMT kernel
  timeout 600 python bench.py --steps 20 --warmup 5 --simt-coarse --no-e2e --no-cpu-baseline --no-gpu-baselines --no-next-rows > $OUT/${TAG}_bench_simt.json 2> $OUT/${TAG}_bench_simt.err
  python scripts/bench_digest.py $OUT/${TAG}_bench_simt.json
  ;;
sweep)
  Q="--no-cpu-baseline --no-gpu-baselines --no-next-rows --no-sweep --steps 10 --warmup 3"
  timeout 900 python bench.py --config 2c --pairs 8 $Q --no-e2e > $OUT/${TAG}_bench_2c_p8.json 2> $OUT/${TAG}_bench_2c_p8.err; python scripts/bench_digest.py $OUT/${TAG}_bench_2c_p8.json
  timeout 900 python bench.py --config 2c --pairs 1 $Q > $OUT/${TAG}_bench_2c_p1.json 2> $OUT/${TAG}_bench_2c_p1.err; python scripts/bench_digest.py $OUT/${TAG}_bench_2c_p1.json
  timeout 900 python bench.py --config indoor --size 640x480 --pairs 4 $Q > $OUT/${TAG}_bench_indoor_p4.json 2> $OUT/${TAG}_bench_indoor_p4.err; python scripts/bench_digest.py $OUT/${TAG}_bench_indoor_p4.json
  for S in 512 832 1024 1152; do
    timeout 900 python bench.py --size $S --global-pairs 16 $Q --no-e2e > $OUT/${TAG}_bench_s${S}_g16.json 2> $OUT/${TAG}_bench_s${S}_g16.err; python scripts/bench_digest.py $OUT/${TAG}_bench_s${S}_g16.json
  done
  ;;
ncu)
  export CASMTR_OVERLAP=0
  BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baselines --no-graph --no-sweep --no-next-rows ${NCU_BENCH_ARGS}"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baselines --no-graph --no-sweep --no-next-rows ${NCU_BENCH_ARGS} > $OUT/${TAG}_ncu_launch.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pool2_tokens|coarse_prep|qtatt_coarse|quad_cta|quad_attention_kernel' -c 5 -f -o $OUT/${TAG}_qtatt $BENCH > $OUT/${TAG}_ncu_a.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cascade_att_tile|quad_attention_list|cascade_match|extract_|fine_match' -c 7 -f -o $OUT/${TAG}_cascade $BENCH > $OUT/${TAG}_ncu_b.log 2>&1
  for r in qtatt cascade; do
    ncu -i $OUT/${TAG}_$r.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_raw.csv 2>/dev/null
  done
  ;;
esac
done
ls -la $OUT | tail -30; du -sh $OUT
