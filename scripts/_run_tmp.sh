timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests -m gpu -x -q -k "not 10816" > gpurun_out/r01z_memcheck.log 2>&1; echo "memcheck exit $?"
tail -6 gpurun_out/r01z_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 30 python -m pytest tests/test_gpu_qtatt.py tests/test_gpu_matching.py tests/test_gpu_golden.py -m gpu -x -q -k "not 104 and not 208" > gpurun_out/r01z_racecheck.log 2>&1; echo "racecheck exit $?"
tail -30 gpurun_out/r01z_racecheck.log | cut -c1-220
