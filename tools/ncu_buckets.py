"""Bucket ncu source-page samples between synchronisation instructions (address order)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
cur = 0; total = 0
keys = ('SYNCS.ARRIVE.TRANS64.A1T0', 'TRYWAIT', 'BAR.SYNC', 'EXIT')
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip(); n = int(r[ix['# Samples']] or 0)
    cur += n; total += n
    if any(k in src for k in keys):
        if cur >= int(sys.argv[2]) if len(sys.argv) > 2 else 300: print(f"{cur:7d}  ..{src[:70]}")
        cur = 0
print("total", total)
