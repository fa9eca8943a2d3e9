"""Kernel time of sf2d reference order for nf frames and k chunk splits (AMEPB_SF2D_SPLIT)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = {}
for nf in (1, 2, 3, 4, 5, 6, 7, 8, 9, 12):
    for k in (1, 2, 3, 4, 6, 8):
        env = dict(os.environ, AMEPB_SF2D_SPLIT=str(k))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_sq.py"), "sf2d_c64", str(nf)], env=env, capture_output=True, text=True).stdout
        ms = float(out.strip().split()[-1])
        rows[(nf, k)] = ms / nf
    print(nf, {k: round(rows[(nf, k)], 2) for k in (1, 2, 3, 4, 6, 8)}, flush=True)
