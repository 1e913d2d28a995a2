"""Summarise the source page of an ncu report: python scripts/ncu_stalls.py <src.csv> [topN]
(src.csv from: ncu -i X.ncu-rep --page source --csv --print-source sass)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# the file may hold several kernels; split on header lines
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        blocks.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
b = blocks[0]
hdr, data = b["hdr"], b["data"]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ci["# Samples"]]) for r in data)
print(b["name"][:80], "instructions", len(data), "samples", tot)
agg = {h: sum(int(r[ci[h]]) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]]))[:topn]
for i in sorted(top):
    r = data[i]
    st = {h[6:]: int(r[ci[h]]) for h in stall_cols if int(r[ci[h]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[ci['# Samples']]):6d} {r[ci['Instructions Executed']]:>10s}  {r[ci['Source']].strip()[:64]:64s} {st}")
