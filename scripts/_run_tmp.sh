timeout 900 python -m pytest tests/test_gpu_attention_layers.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python scripts/_probe_tokens.py 2>&1 | tail -8
