timeout 300 python -m pytest tests/test_kquants_gpu.py tests/test_gguf_gpu.py -x -q 2>&1 | tail -15
