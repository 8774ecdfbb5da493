timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_d.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'])
print(d['roofline'])
print(d['roofline_int8'])
for r in d['gemm_shapes']['bf16'][:6]: print(r)
print({k:v['ms_per_step'] for k,v in d['kernels'].items()})
print(d['prefix_lm'])
PY
