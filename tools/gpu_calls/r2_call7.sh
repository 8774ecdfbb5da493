export WD=llamax_b200/csrc/libllamax_b200_wd.so
LLAMAX_B200_LIB=$WD timeout 300 python -m pytest tests/test_attention_gpu.py -x -q > gpurun_out/r2_attn_wd.log 2>&1; rc=$?; echo "attn wd rc=$rc"; tail -3 gpurun_out/r2_attn_wd.log
if [ $rc -ne 0 ]; then exit 1; fi
echo "== v2"; LLAMAX_ATTN_FWD=2 timeout 120 python tools/attn_fwd_perf.py
echo "== v4 token"; timeout 120 python tools/attn_fwd_perf.py
echo "== v4 no token"; LLAMAX_ATTN_STAGGER=0 timeout 120 python tools/attn_fwd_perf.py
timeout 120 python tools/attn_trace.py fwd > gpurun_out/r2_trace_v4t.log 2>&1; cat gpurun_out/r2_trace_v4t.log
