# final-code check: full GPU suite, smoke, default bench line
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
timeout 900 python bench.py > gpurun_out/r2_bench_1gpu_s3.json 2> gpurun_out/r2_bench_1gpu_s3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_1gpu_s3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_1gpu_s3.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","roofline","clocks")})
print(d.get("prefix_lm")); print(d.get("kernels"))
PY
