timeout 900 python -m pytest tests/test_elementwise_gpu.py tests/test_block_gpu.py -m gpu -x -q 2>&1 | tail -6
