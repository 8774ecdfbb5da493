timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_block_gpu.py tests/test_cabi.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_ab_attnside.json 2> gpurun_out/r2_ab_attnside.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_ab_attnside.json").read().strip().splitlines()[-1])
k=d["kernels"]
print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], {n:k[n]["ms_per_step"] for n in ("attn_bwd","attn_fwd","bf16_gemm","int8_gemm")})
PY
