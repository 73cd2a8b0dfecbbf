timeout 300 python -m pytest tests/test_gpu_coarse_match.py -x -q 2>&1 | tail -15
