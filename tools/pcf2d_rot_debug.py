"""Rotated pcf2d: float32-certified path vs float64-only path vs the oracle on the dense test case."""
import os, sys, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import amep_oracle as orc
from amep_b200 import spatialcor as sc, _core
for bd in (np.float32, np.float64):
    rng = np.random.default_rng(zlib.crc32(f"pcf2d-rot-dense-{np.dtype(bd).name}".encode()))
    n = 8000
    box = np.array([[-5, 55], [10, 70], [-0.5, 0.5]], dtype=bd)
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, 0] = rng.uniform(-5, 55, n).astype(np.float32)
    c[:, 1] = rng.uniform(10, 70, n).astype(np.float32)
    psi = np.array([-0.35, -0.61])
    want, xb, yb = orc.pcf2d_counts(c, box, None, 400, 380, 20.0, psi)
    res = {}
    for mode in ("fast", "slow"):
        if mode == "slow": os.environ["AMEPB_PCF2D_NO_F32ROT"] = "1"
        else: os.environ.pop("AMEPB_PCF2D_NO_F32ROT", None)
        ang = np.arccos(psi[0] / np.sqrt(psi[0] ** 2 + psi[1] ** 2)); ang = 2 * np.pi - ang
        cen = np.mean(box, axis=1)
        r = np.array([[1.0, np.cos(-ang), np.sin(-ang)]])
        got = _core.pcf2d_counts(torch.from_numpy(c).cuda(), None, cen.astype(np.float64)[None], cen.dtype == np.float32, r, xb, yb)[0].cpu().numpy()
        res[mode] = got
        d = got - want
        print(np.dtype(bd).name, mode, "sum got", got.sum(), "want", want.sum(), "bins differing", np.count_nonzero(d), "abs diff", np.abs(d).sum(), "net", d.sum())
    print(" fast vs slow bins differing", np.count_nonzero(res["fast"] - res["slow"]))
