timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -s 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_2gpu.json"))
print(d['value'], d['ms_per_step'], d.get('dp'))
print(d['prefix_lm']['value'], d['prefix_lm']['ms_per_step'], d['prefix_lm'].get('dp'))
PY
tail -5 gpurun_out/r2_bench_2gpu.err
