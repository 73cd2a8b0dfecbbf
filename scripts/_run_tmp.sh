timeout 900 python -m pytest tests/test_gpu_attention_layers.py tests/test_gpu_qtatt.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -15
