# 4-GPU record at HEAD under the driver's torchrun line
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_4gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks']); print(d['dp']); print(d['prefix_lm']['value'], d['prefix_lm']['ms_per_step'], d['prefix_lm']['dp'])
PY
