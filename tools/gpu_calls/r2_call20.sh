echo "== v4 token + alu pack"; LLAMAX_ATTN_ALU_PACK=1 timeout 300 python tools/attn_fwd_perf.py
LLAMAX_ATTN_ALU_PACK=1 timeout 120 python tools/attn_trace.py fwd 2>&1 | head -22
