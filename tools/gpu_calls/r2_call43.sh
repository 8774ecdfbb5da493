timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final4.json 2> gpurun_out/r2_bench_1gpu_final4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu_final4.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","clocks")})
p=d["prefix_lm"]; print({k:p[k] for k in ("value","ms_per_step","attn_fwd_tflops","attn_bwd_tflops","attn_ms_per_step")})
print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline_int8"]["achieved"], d["roofline_int8"]["frac"], d.get("gpu_launches"))
for k,v in d["kernels"].items(): print(k,v)
PY
