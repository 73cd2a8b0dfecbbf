timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r01i_bench.json 2>gpurun_out/r01i_bench.err; tail -3 gpurun_out/r01i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r01i_bench.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['matches_per_step'])
for r in d['breakdown']: print(r)
PY
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'transpose_vec|qtatt_coarse|quad_cta' -c 4 -f -o gpurun_out/r01i_qtatt $BENCH > gpurun_out/r01i_ncu_a.log 2>&1
ncu -i gpurun_out/r01i_qtatt.ncu-rep --page raw --csv > gpurun_out/r01i_qtatt_raw.csv 2>/dev/null
