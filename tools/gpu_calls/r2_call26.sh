timeout 600 python -m pytest tests/test_block_gpu.py -m gpu -x -q -k "weight_cache" 2>&1 | tail -4
LLAMAX_WEIGHT_CACHE=int8 timeout 900 python bench.py --workload text --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_bench_cache_int8.json 2> gpurun_out/r2_bench_cache_int8.err; echo "rc=$?"
timeout 900 python bench.py --workload text --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_bench_cache_auto.json 2> gpurun_out/r2_bench_cache_auto.err; echo "rc=$?"
python - <<'PY'
import json
for f in ["r2_bench_cache_int8","r2_bench_cache_auto"]:
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, d['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['kernels'].get('dequant_weight'))
PY
