timeout 600 python -m pytest tests/test_elementwise_gpu.py -m gpu -x -q 2>&1 | tail -3
EW_ONLY=copy,rmsnorm_fwd,rowquant timeout 300 python tools/ew_sustained.py 12 | grep -E "copy 134|rmsnorm|rowquant|load"
timeout 300 python tools/ew_perf.py 2>/dev/null | grep -E "rmsnorm_fwd|rowquant"
