run() { name=$1; shift; env "$@" timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run r2_ab_as_new X=1
run r2_ab_as_prev LLAMAX_B200_LIB=llamax_b200/csrc/libllamax_b200_prev.so
run r2_ab_as_new2 X=1
run r2_ab_as_prev2 LLAMAX_B200_LIB=llamax_b200/csrc/libllamax_b200_prev.so
python - <<'PY'
import json
for f in ["r2_ab_as_new","r2_ab_as_prev","r2_ab_as_new2","r2_ab_as_prev2"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], {n:k[n]["ms_per_step"] for n in ("attn_bwd","attn_fwd","bf16_gemm","int8_gemm")})
PY
