"""Summarise an ncu source page (ncu -i rep --page source --csv): hottest SASS lines by stall samples.
   python scripts/ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
tot_inst = sum(int(r[ix["Instructions Executed"]] or 0) for r in data)
print("total samples", tot, "warp-instructions executed", tot_inst, "static instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print("stall mix:", {k: round(100 * v / max(tot, 1), 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    main = max(stalls, key=lambda k: int(r[ix[k]] or 0))
    print(f"{i:5d} {100*s/tot:5.1f}%  exec={r[ix['Instructions Executed']]:>8s} {main:18s} {r[ix['Source']].strip()[:90]}")
