echo "== v4 (default)"; timeout 300 python tools/attn_fwd_perf.py bwd sdpa
echo "== v2"; LLAMAX_ATTN_FWD=2 timeout 300 python tools/attn_fwd_perf.py
