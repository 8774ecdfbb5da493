timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py tests/test_cabi.py -m gpu -x -q 2>&1 | tail -6
run() { name=$1; shift; env "$@" timeout 900 python bench.py --workload text --steps 10 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -2 gpurun_out/$name.err; }
run r2_ab_ropeepi_0 X=1
run r2_ab_ropeepi_1 LLAMAX_ROPE_EPILOGUE=1
run r2_ab_ropeepi_0b X=1
run r2_ab_ropeepi_1b LLAMAX_ROPE_EPILOGUE=1
python - <<'PY'
import json
for f in ["r2_ab_ropeepi_0","r2_ab_ropeepi_1","r2_ab_ropeepi_0b","r2_ab_ropeepi_1b"]:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    q=[(r["ms_per_step"], r["achieved"]) for r in d["gemm_shapes"]["int8"] if "N=6144" in r["shape"]]
    print(f, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], "rope", k.get("rope",{}).get("ms_per_step"), "int8", k["int8_gemm"]["ms_per_step"], "qkv", q, "loss", d.get("loss"))
PY
