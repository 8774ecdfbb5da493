timeout 100 python -m pytest tests/test_block_gpu.py -k "compaction or tiny_model or trainable" -x -q 2>&1 | tail -3
LLAMAX_ROW_WPR=1 timeout 100 python -m pytest tests/test_elementwise_gpu.py -x -q 2>&1 | tail -2
run() { n=$1; shift; env "$@" timeout 200 python bench.py --steps 8 --warmup 3 --workload text --no-cpu-baseline --no-int8-peak > gpurun_out/r2_ab_$n.json 2> gpurun_out/r2_ab_$n.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_ab_$n.json").read().strip().splitlines()[-1])
k=d["kernels"]
print("$n", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], "bf16", k["bf16_gemm"]["ms_per_step"], "rmsfwd", k["rmsnorm_fwd"]["ms_per_step"], "rowq", k["rowquant"]["ms_per_step"], "ce", k["cross_entropy"]["ms_per_step"])
PY
}
run base A=1
run nocompact LLAMAX_LM_COMPACT=0
run wpr LLAMAX_ROW_WPR=1
run base2 A=1
