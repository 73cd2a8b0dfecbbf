#!/bin/bash
# ncu captures of the main kernels, kept small enough to travel back (gpurun_out/ <= 64 MiB):
#   gpurun --timeout 900 -- 'bash scripts/gpu_ncu.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"
# launch list of two whole steps (after 3 warm-up steps = 183 launches of ours; torch's own kernels are in the list too)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_launch.log 2>&1
# full sets: first QTAttB call (layout, coarse, mid, last), one cascade call, the match kernel
ncu --set full --clock-control none --import-source on -k regex:'transpose_jobs|qtatt_coarse|quad_attention' -c 4 -f -o $OUT/${TAG}_qtatt $BENCH > $OUT/${TAG}_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'transpose_jobs|quad_attention' -s 36 -c 2 -f -o $OUT/${TAG}_cascade $BENCH > $OUT/${TAG}_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'cascade_match|extract_|fine_match' -c 5 -f -o $OUT/${TAG}_match $BENCH > $OUT/${TAG}_ncu_c.log 2>&1
for r in qtatt cascade match; do
  ncu -i $OUT/${TAG}_$r.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_raw.csv 2>/dev/null
done
ls -la $OUT; du -sh $OUT
