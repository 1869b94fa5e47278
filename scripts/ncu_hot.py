"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples, with stall reasons."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[idx["# Samples"]])
    except ValueError: continue
    tot += n
    st = {h: int(r[idx[h]] or 0) for h in stall_cols}
    data.append((n, r[idx["Address"]], r[idx["Source"]], st, r[idx["Instructions Executed"]]))
print("total samples", tot, "instructions", len(data))
agg = {}
for n, a, s, st, ie in data:
    for h, v in st.items(): agg[h] = agg.get(h, 0) + v
print("stall totals:", sorted(((v, h) for h, v in agg.items() if v), reverse=True)[:10])
# order-preserving listing of hot instructions
for n, a, s, st, ie in sorted(data, key=lambda x: -x[0])[:topn]:
    top = sorted(((v, h.replace("stall_", "")) for h, v in st.items() if v), reverse=True)[:3]
    print(f"{n:6d} {100*n/tot:5.1f}%  exec={ie:>8s}  {s[:90]:90s} {top}")
