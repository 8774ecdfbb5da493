timeout 900 python -m pytest tests/test_block_gpu.py tests/test_gemm_gpu.py tests/test_seams_gpu.py -m gpu -x -q 2>&1 | tail -4
run() { name=$1; shift; env "$@" timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; }
run r2_ab_lora_new X=1
run r2_ab_lora_old LLAMAX_LORA_GROUP_PAIR=0 LLAMAX_LORA_PAIR_SPLITS=3
run r2_ab_lora_new2 X=1
python - <<'PY'
import json
for f in ["r2_ab_lora_new","r2_ab_lora_old","r2_ab_lora_new2"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], {n:(k[n]["ms_per_step"], k[n].get("launches_per_step")) for n in ("lora_wgrad","batched_copy","rmsnorm_bwd","bf16_gemm")})
PY
