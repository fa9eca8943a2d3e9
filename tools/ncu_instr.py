"""Executed-instruction profile of an .ncu-rep (source page): SASS listing with per-instruction
warp-level execution counts, so that loop bodies and their trip counts can be read off."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
tot = 0
lines = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ci["Instructions Executed"]])
    except ValueError: continue
    tot += n
    lines.append((r[ci["Address"]], r[ci["Source"]].strip(), n, int(r[ci["# Samples"]] or 0)))
print("total warp instructions", tot)
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
for i, (a, s, n, smp) in enumerate(lines):
    if n >= thr * tot:
        print(f"{i:5d} {n:11d} {n/tot:6.2%} {smp:6d}  {s[:90]}")
