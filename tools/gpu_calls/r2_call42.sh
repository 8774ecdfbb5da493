# 2 GPUs: NCCL DP parity test + the 2-GPU bench line with the final kernels
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/r2_bench_2gpu_final2.json 2> gpurun_out/r2_bench_2gpu_final2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_2gpu_final2.json").read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","n_gpus","e2e","dp")})
print(d["prefix_lm"]["value"], d["prefix_lm"]["ms_per_step"], d["prefix_lm"].get("dp"))
PY
