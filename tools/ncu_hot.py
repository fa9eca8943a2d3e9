"""Top stalled SASS instructions of an .ncu-rep (source page), with the dominant stall reason."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: s = int(r[ci["# Samples"]])
    except ValueError: continue
    data.append((s, r))
tot = sum(s for s, _ in data)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("total samples", tot)
agg = {c: sum(int(r[ci[c]] or 0) for _, r in data) for c in stall_cols}
print("by reason:", ", ".join(f"{k[6:]}={v/tot:.1%}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for s, r in sorted(data, key=lambda x: -x[0])[:n]:
    top = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:2]
    print(f"{s:7d} {s/tot:6.1%}  {r[ci['Source']].strip()[:70]:70s} {top[0][1]}:{top[0][0]} {top[1][1]}:{top[1][0]}")
