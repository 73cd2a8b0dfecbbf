timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r01h_bench.json 2>gpurun_out/r01h_bench.err; tail -3 gpurun_out/r01h_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01h_bench.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['matches_per_step'])
for r in d['breakdown']: print(r)
PY
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'cascade_match_tile|cascade_match_cell|cascade_att_tile' -s 3 -c 3 -f -o gpurun_out/r01h_match $BENCH > gpurun_out/r01h_ncu_b.log 2>&1
ncu -i gpurun_out/r01h_match.ncu-rep --page raw --csv > gpurun_out/r01h_match_raw.csv 2>/dev/null
