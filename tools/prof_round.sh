set -x
python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_e.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_e1.log 2>&1
REPS=2 python tools/prof_kernels.py > gpurun_out/plain_e2.log 2>&1 &&
REPS=2 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'gemm_kernel|attn_fwd_kernel|attn_bwd_kernel|lora_bwd_pair_kernel|lora_wgrad_tc_kernel|rmsnorm_fwd_kernel|rmsnorm_bwd_kernel|swiglu_fwd_kernel' -s 11 -c 11 -o gpurun_out/prof_r1e python tools/prof_kernels.py > gpurun_out/ncu_e2.log 2>&1
tail -3 gpurun_out/ncu_e1.log gpurun_out/ncu_e2.log; ls -la gpurun_out/prof_r1e* gpurun_out/launches_r1e.csv
