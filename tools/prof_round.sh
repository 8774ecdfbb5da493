# ncu recipe of the round (run under gpurun, one GPU): launch list of a 2-layer bench + full-set capture of the hot kernels
set -x
python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain_e.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --layers 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_e1.log 2>&1
REPS=1 python tools/prof_kernels.py > gpurun_out/plain_e2.log 2>&1 &&
REPS=1 ncu --set full --clock-control none --import-source on --kernel-name-base function -k regex:'^gemm_kernel|attn_fwd2_kernel|attn_bwd_kernel|lora_bwd_pair_kernel|lora_wgrad_tc_kernel|rmsnorm_fwd|rmsnorm_bwd_kernel|swiglu_fwd_kernel|rowquant' -c 14 -o gpurun_out/prof_r1i python tools/prof_kernels.py > gpurun_out/ncu_e2.log 2>&1
ls -la gpurun_out/prof_r1i* gpurun_out/launches_r1i.csv
