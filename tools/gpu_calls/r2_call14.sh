timeout 120 python tools/gemm_vs_cublas.py ncu > gpurun_out/plain14.log 2>&1 &&
timeout 900 ncu --set full --clock-control none -k regex:'gemm_kernel|nvjet|cutlass|sm100|gemm' -s 2 -c 2 -o gpurun_out/prof_r2_gemm_cublas python tools/gemm_vs_cublas.py ncu > gpurun_out/ncu14.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu14.log; ls -la gpurun_out/prof_r2_gemm_cublas*
