echo "== default"; timeout 300 python tools/ew_sustained.py 12
echo "== LLAMAX_ROW_WPR=1"; LLAMAX_ROW_WPR=1 timeout 300 python tools/ew_sustained.py 8 | grep -E "rmsnorm_fwd|rowquant|load"
echo "== LLAMAX_ROW_RING_CTAS=4"; LLAMAX_ROW_RING_CTAS=4 timeout 300 python tools/ew_sustained.py 8 | grep -E "rmsnorm_fwd|rowquant|load"
echo "== LLAMAX_ROW_RING=0"; LLAMAX_ROW_RING=0 timeout 300 python tools/ew_sustained.py 8 | grep -E "rmsnorm_fwd|rowquant|load"
