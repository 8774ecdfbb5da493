# 8-GPU record at HEAD under the driver's torchrun line
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 300 gpurun_out/r2_bench_8gpu.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_8gpu.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks']); print(d['dp']); print(d['prefix_lm']['value'], d['prefix_lm']['ms_per_step'], d['prefix_lm']['dp'])
PY
