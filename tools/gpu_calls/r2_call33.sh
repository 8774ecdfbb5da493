REPS=2 python tools/prof_rows.py > gpurun_out/plain_rows.log 2>&1 &&
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'row_wpr_kernel|rmsnorm_bwd_ring_kernel|swiglu_fwd_ring_kernel' -s 4 -c 4 -o gpurun_out/prof_rows python tools/prof_rows.py > gpurun_out/ncu_rows.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_rows.log; ls -la gpurun_out/prof_rows.ncu-rep
