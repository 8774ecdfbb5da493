# ncu recipe, final kernels of round 2 (run under gpurun, one GPU): launch list of a 2-layer bench + full-set capture of the hot kernels
python bench.py --workload text --layers 2 --steps 1 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/plain_r2e.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --workload text --layers 2 --steps 1 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/ncu_r2e.log 2>&1
echo "launch list rc=$?"
REPS=1 python tools/prof_kernels.py > gpurun_out/plain_r2f.log 2>&1 &&
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'^gemm_kernel|gemm_wide_kernel|attn_fwd._kernel|attn_bwd_kernel|lora_bwd_pair_kernel|lora_wgrad_tc_kernel|row_wpr_kernel|rmsnorm_bwd_ring_kernel|swiglu_fwd_ring_kernel|rope_kernel' -c 18 -o gpurun_out/prof_r2c python tools/prof_kernels.py > gpurun_out/ncu_r2f.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/prof_r2c.ncu-rep gpurun_out/launches_r2c.csv
