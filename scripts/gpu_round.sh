#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), ncu launch list, ncu --set full of the main kernels.
# Run as:  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
# Everything lands in gpurun_out/ (kept under 64 MiB); scripts/ncu_*.py turn the exports into profiles/ summaries.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)" >> $OUT/${TAG}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json | cut -c1-600
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref exit $?"
cat $OUT/${TAG}_bench_ref.json | cut -c1-300
# per-kernel counters: one transpose launch per call (CASMTR_OVERLAP=0) keeps the -s / -c launch arithmetic below simple; the
# launch list above is taken with the library's defaults
export CASMTR_OVERLAP=0
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'transpose_vec|qtatt_coarse|quad_cta|quad_attention_kernel' -c 4 -f -o $OUT/${TAG}_qtatt $BENCH > $OUT/${TAG}_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'transpose_vec|cascade_att_tile|quad_attention_list' -s 12 -c 3 -f -o $OUT/${TAG}_cascade $BENCH > $OUT/${TAG}_ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cascade_match|extract_|fine_match' -c 6 -f -o $OUT/${TAG}_match $BENCH > $OUT/${TAG}_ncu_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'coarse_rowstats|pool2_tokens|fine_window_gather|tf32_residual' -c 4 -f -o $OUT/${TAG}_widen $BENCH > $OUT/${TAG}_ncu_d.log 2>&1
for k in relative_pe_kernel score5d_bwd_kernel value_agg_bwd_kernel score3d_bwd_kernel; do       # one launch of each (they are timed in loops)
  timeout 600 ncu --set full --clock-control none --import-source on -k $k -c 1 -f -o $OUT/${TAG}_w2_$k $BENCH > $OUT/${TAG}_ncu_e_$k.log 2>&1
done
for r in qtatt cascade match widen w2_relative_pe_kernel w2_score5d_bwd_kernel w2_value_agg_bwd_kernel w2_score3d_bwd_kernel; do
  ncu -i $OUT/${TAG}_$r.ncu-rep --page raw --csv > $OUT/${TAG}_${r}_raw.csv 2>/dev/null
done
ls -la $OUT; du -sh $OUT
