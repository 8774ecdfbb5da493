timeout 600 python -m pytest tests/test_elementwise_gpu.py tests/test_block_gpu.py tests/test_gemm_gpu.py -m gpu -x -q 2>&1 | tail -5
echo "== new defaults"; timeout 300 python tools/ew_sustained.py 12 | grep -E "copy|rmsnorm|rowquant|swiglu|rope|load"
echo "== isolated"; timeout 300 python tools/ew_perf.py 2>/dev/null | head -8
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'row_wpr_kernel|rmsnorm_bwd_ring_kernel|swiglu_fwd_ring_kernel' -s 4 -c 4 -o gpurun_out/prof_rows2 python tools/prof_rows.py > gpurun_out/ncu_rows2.log 2>&1
echo "rc=$?"
