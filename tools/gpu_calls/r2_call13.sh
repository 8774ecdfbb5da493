timeout 600 python tools/gemm_vs_cublas.py 2>&1 | tee gpurun_out/r2_gemm_vs_cublas.log
timeout 120 python tools/gemm_vs_cublas.py ncu > gpurun_out/plain13.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k regex:'gemm_kernel|nvjet|cutlass|gemm' -s 4 -c 4 -o gpurun_out/prof_r2_gemm_cublas python tools/gemm_vs_cublas.py ncu > gpurun_out/ncu13.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu13.log
