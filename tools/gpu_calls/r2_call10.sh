timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest10.log
timeout 600 python -m pytest tests/test_block_gpu.py -m gpu -q -s -k "matches_oracle or rank16 or partial or ragged" > gpurun_out/r2_parity_tables.log 2>&1
timeout 900 python bench.py --workload text --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak --weight-only > gpurun_out/r2_bench_weight_only.json 2> gpurun_out/r2_bench_weight_only.err; echo "wo rc=$?"
LLAMAX_WEIGHT_CACHE=0 timeout 900 python bench.py --workload text --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_bench_cache0.json 2> gpurun_out/r2_bench_cache0.err; echo "c0 rc=$?"
timeout 900 python bench.py --workload text --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "c rc=$?"
python - <<'PY'
import json
for f in ["r2_bench_weight_only","r2_bench_cache0","r2_bench_c"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], {k:v['ms_per_step'] for k,v in d['kernels'].items()})
    except Exception as e:
        print(f, "failed", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv
