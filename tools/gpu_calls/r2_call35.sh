timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_ab_rows_new.json 2> gpurun_out/r2_ab_rows_new.err; echo "rc=$?"
LLAMAX_ROW_WPR=0 LLAMAX_RMSNORM_BWD_RING=0 LLAMAX_SWIGLU_RING=0 timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_ab_rows_old.json 2> gpurun_out/r2_ab_rows_old.err; echo "rc=$?"
timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_ab_rows_new2.json 2> gpurun_out/r2_ab_rows_new2.err; echo "rc=$?"
python - <<'PY'
import json
for f in ["r2_ab_rows_new","r2_ab_rows_old","r2_ab_rows_new2"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], {n:(k[n]["ms_per_step"], k[n].get("achieved_gbs")) for n in ("rmsnorm_fwd","rmsnorm_bwd","swiglu_fwd","rope","rowquant","lora_wgrad")})
PY
