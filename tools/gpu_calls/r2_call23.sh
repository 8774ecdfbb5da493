timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
/usr/bin/time -v timeout 1500 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json | cut -c1-400
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_e.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e'])
print(d['roofline']); print(d['roofline_int8'])
for r in d['gemm_shapes']['bf16'][:6]: print(r)
print({k:(v['ms_per_step'], v.get('achieved_gbs') or v.get('achieved_tflops')) for k,v in d['kernels'].items()})
print(d['prefix_lm']); print(d['cpu_baseline']); print(d['gpu_launches'])
PY
