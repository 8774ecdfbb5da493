export WD=llamax_b200/csrc/libllamax_b200_wd.so
LLAMAX_B200_LIB=$WD timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "wide" > gpurun_out/r2_gemm_wd.log 2>&1; rc=$?; echo "gemm wd rc=$rc"; tail -15 gpurun_out/r2_gemm_wd.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/gemm_vs_cublas.py 2>&1 | tee gpurun_out/r2_gemm_vs_cublas_wide.log
