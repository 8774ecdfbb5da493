timeout 600 python -m pytest tests/test_elementwise_gpu.py tests/test_block_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "== new defaults"; timeout 300 python tools/ew_sustained.py 12
echo "== old kernels (LLAMAX_ROW_WPR=0 LLAMAX_RMSNORM_BWD_RING=0 LLAMAX_SWIGLU_RING=0)"; LLAMAX_ROW_WPR=0 LLAMAX_RMSNORM_BWD_RING=0 LLAMAX_SWIGLU_RING=0 timeout 300 python tools/ew_sustained.py 12 | grep -E "rmsnorm|rowquant|swiglu|rope|load"
echo "== isolated"; timeout 300 python tools/ew_perf.py | head -12
