timeout 900 python -m pytest tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -15
