# 2-GPU check of HEAD: NCCL DP parity test + the driver's torchrun launch of bench.py at N=2 (both arms)
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2_bench_2gpu_reentry.json 2> gpurun_out/r2_bench_2gpu_reentry.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r2_bench_2gpu_reentry.json | head -c 1500
