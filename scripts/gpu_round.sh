#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list, ncu --set full of one whole step.
# Run as:  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
# Everything lands in gpurun_out/; summaries worth keeping are copied into profiles/ by scripts/summarize_ncu.py.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json
if [ "${SKIP_REF:-0}" != "1" ]; then
  python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
  cat $OUT/${TAG}_bench_ref.json
fi
if [ "${SKIP_NCU:-0}" != "1" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
  ncu --set full --clock-control none --import-source on -c 75 -f -o $OUT/${TAG}_step \
      python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
  ls -la $OUT
fi
