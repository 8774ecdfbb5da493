set -x
export WD=llamax_b200/csrc/libllamax_b200_wd.so
LLAMAX_B200_LIB=$WD timeout 300 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/r2_attn_wd.log 2>&1; rc=$?; echo "attn wd rc=$rc"; tail -5 gpurun_out/r2_attn_wd.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest3.log
for v in 2 4 2 4; do LLAMAX_ATTN_FWD=$v timeout 120 python tools/attn_fwd_perf.py > gpurun_out/r2_fwdperf_v$v.log 2>&1; echo "== fwd v$v"; cat gpurun_out/r2_fwdperf_v$v.log; done
