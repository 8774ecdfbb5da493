"""Summarise an `ncu --page source --csv` dump: top SASS instructions by warp-stall samples.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-id :::N > src.csv ; python tools/ncu_hot.py src.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print(rows[0][1][:100] if len(rows[0]) > 1 else "", "| instructions", len(data), "| samples", tot)
agg = {h: sum(int(r[hdr.index(h)]) for r in data) for h in stalls}
print("stall mix:", {k: f"{100*v/max(tot,1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
    st = {h: int(r[hdr.index(h)]) for h in stalls}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{int(r[si]):6d} {100*int(r[si])/tot:5.1f}% exec={r[ie]:>8} {r[src].strip()[:66]:66s} {main}")
