timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_final3.json 2> gpurun_out/r2_bench_1gpu_final3.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu_final3.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","clocks")})
p=d["prefix_lm"]; print({k:p[k] for k in ("value","ms_per_step","e2e","attn_fwd_tflops","attn_bwd_tflops","attn_ms_per_step")})
print(d["roofline"]["achieved"], d["roofline"]["frac"], d["roofline_int8"]["achieved"], d["roofline_int8"]["frac"])
for k,v in d["kernels"].items(): print(k,v)
PY
