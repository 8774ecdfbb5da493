set -x
export WD=llamax_b200/csrc/libllamax_b200_wd.so
LLAMAX_B200_LIB=$WD timeout 300 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/r2_attn_wd.log 2>&1; rc=$?; echo "attn wd rc=$rc"; tail -5 gpurun_out/r2_attn_wd.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest2.log
for v in 2 3 2 3; do LLAMAX_ATTN_FWD=$v timeout 120 python tools/attn_fwd_perf.py > gpurun_out/r2_fwdperf_v$v.log 2>&1; echo "== fwd v$v"; cat gpurun_out/r2_fwdperf_v$v.log; done
timeout 600 python bench.py --workload text --steps 5 --warmup 3 --no-cpu-baseline --no-int8-peak > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_b.json'))
print(d['value'], d['ms_per_step'], d['clocks'])
for r in d['gemm_shapes']['bf16'][:6]: print(r)
for k,v in d['kernels'].items(): print(k, v)
PY
