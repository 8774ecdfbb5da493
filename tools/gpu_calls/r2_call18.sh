timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_block_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/attn_fwd_perf.py bwd
