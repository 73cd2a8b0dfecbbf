BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cascade_att_tile|cascade_match_tile' -c 2 -f -o gpurun_out/r01r_tiles $BENCH > gpurun_out/r01r_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'quad_cta|quad_attention_kernel' -c 2 -f -o gpurun_out/r01r_fine $BENCH > gpurun_out/r01r_ncu2.log 2>&1
for r in tiles fine; do ncu -i gpurun_out/r01r_$r.ncu-rep --page raw --csv > gpurun_out/r01r_${r}_raw.csv 2>/dev/null; done
