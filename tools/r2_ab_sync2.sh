LLAMAX_B200_LIB=llamax_b200/csrc/libllamax_b200_wd.so timeout 100 python -m pytest tests/test_gemm_gpu.py -k "swiglu" -x -q 2>&1 | tail -2
run() { n=$1; shift; env "$@" timeout 200 python bench.py --steps 8 --warmup 3 --workload text --no-cpu-baseline --no-int8-peak > gpurun_out/r2_ab_$n.json 2> gpurun_out/r2_ab_$n.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r2_ab_$n.json").read().strip().splitlines()[-1])
k=d["kernels"]
sw=[r for r in d["gemm_shapes"]["bf16"] if r["shape"]=="[M=16384,N=14336,K=4096]"][0]
print("$n", d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], "bf16", k["bf16_gemm"]["ms_per_step"], "w2bwd", sw["ms_per_step"], sw["achieved"])
PY
}
run swi2 A=1
run swi0 LLAMAX_GEMM_SYNC_EVERY=0
run swi2b A=1
run swi0b LLAMAX_GEMM_SYNC_EVERY=0
