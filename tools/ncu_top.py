"""Summarise an ncu source page CSV: total samples by stall reason and the top-N SASS lines."""
import csv, sys
from collections import Counter
path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = Counter(); lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    n = int(r[ix["# Samples"]] or 0)
    for s in stall_cols:
        tot[s] += int(r[ix[s]] or 0)
    lines.append((n, r[ix["Source"]].strip(), {s: int(r[ix[s]] or 0) for s in stall_cols if int(r[ix[s]] or 0)}))
total = sum(n for n, _, _ in lines)
print("total samples", total)
for s, c in tot.most_common(10): print(f"  {s:28s} {c:8d} {100*c/max(total,1):5.1f}%")
for n, src, st in sorted(lines, key=lambda t: -t[0])[:topn]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{n:7d} {100*n/max(total,1):5.1f}%  {src[:70]:70s} {top}")
