echo "== v2"; LLAMAX_ATTN_FWD=2 timeout 120 python tools/attn_fwd_perf.py bwd
echo "== v4 token"; timeout 120 python tools/attn_fwd_perf.py
echo "== v4 no token"; LLAMAX_ATTN_STAGGER=0 timeout 120 python tools/attn_fwd_perf.py
