timeout 300 python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -2
echo "== v2"; LLAMAX_ATTN_FWD=2 timeout 120 python tools/attn_fwd_perf.py
echo "== v4 stagger"; timeout 120 python tools/attn_fwd_perf.py
echo "== v4 nostagger"; LLAMAX_ATTN_STAGGER=0 timeout 120 python tools/attn_fwd_perf.py
timeout 120 python tools/attn_trace.py fwd > gpurun_out/r2_trace_v4s.log 2>&1; cat gpurun_out/r2_trace_v4s.log
