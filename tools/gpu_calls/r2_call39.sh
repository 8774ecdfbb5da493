run() { name=$1; shift; env "$@" timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run r2_ab_lmh_256 X=1
run r2_ab_lmh_wide LLAMAX_GEMM_WIDE_LMHEAD=1
run r2_ab_lmh_256b X=1
run r2_ab_lmh_wideb LLAMAX_GEMM_WIDE_LMHEAD=1
python - <<'PY'
import json
for f in ["r2_ab_lmh_256","r2_ab_lmh_wide","r2_ab_lmh_256b","r2_ab_lmh_wideb"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], [(r["shape"], r["ms_per_step"], r["achieved"]) for r in d["gemm_shapes"]["bf16"] if "N=128256" in r["shape"]])
PY
