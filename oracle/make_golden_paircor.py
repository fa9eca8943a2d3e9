"""Golden vectors of ``amep.spatialcor.pcf_angle`` and ``amep.spatialcor.spatialcor`` produced by
the UNMODIFIED reference (AMEP 1.3.0).  Run in the build container only (needs /root/reference):

    python oracle/make_golden_paircor.py

pcf_angle fixtures also store the integer pair counts of the chunk worker ``__dhist_angle``
(amep/spatialcor.py:929-1041); spatialcor fixtures the pair counts per distance bin (``norm``) and the
unnormalised sums of the chunk worker ``__scor_chunk`` (amep/spatialcor.py:47-135).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def pcfa_case(amep, name, coords, box, other=None, **kw):
    dh = getattr(amep.spatialcor, "__dhist_angle")
    c0 = coords.copy()
    o0 = None if other is None else other.copy()
    g, R, T = amep.spatialcor.pcf_angle(coords, box, other_coords=other, **kw)
    # raw counts of the worker, same arguments as pcf_angle builds them (amep/spatialcor.py:1244-1283)
    rmax = kw.get("rmax")
    if rmax is None:
        rmax = max(box[:, 1] - box[:, 0]) / 2
    ndb, nab = kw.get("ndbins", 500), kw.get("nabins", 100)
    dbins = np.linspace(0.0, rmax, ndb + 1)
    abins = np.linspace(0.0, 2 * np.pi, nab + 1)
    psi = kw.get("psi")
    angle = 0.0
    if psi is not None:
        angle = np.arccos(np.dot(np.array([1, 0]), psi) / np.sqrt(psi[0] ** 2 + psi[1] ** 2))
        if psi[1] < 0:
            angle = 2 * np.pi - angle
    e = kw.get("e", np.array([1.0, 0.0, 0.0]))
    oc = coords if other is None else other
    counts = dh(0, len(oc), coords, box, oc, dbins, abins, e, angle, other is None, kw.get("pbc", True))
    out = dict(coords=c0, box=box, g=g, R=R, T=T, counts=counts.astype(np.int64), ndbins=ndb, nabins=nab,
               rmax=np.float64(np.nan if kw.get("rmax") is None else kw["rmax"]), pbc=np.bool_(kw.get("pbc", True)),
               e=np.asarray(e))
    if psi is not None:
        out["psi"] = np.asarray(psi)
    if other is not None:
        out["other"] = o0
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: pairs={int(counts.sum())} g.dtype={g.dtype} shape={g.shape} dbins.dtype={dbins.dtype}")


def scor_case(amep, name, coords, box, values, other=None, other_values=None, **kw):
    chunk = getattr(amep.spatialcor, "__scor_chunk")
    c, r = amep.spatialcor.spatialcor(coords, box, values, other_coords=other, other_values=other_values, **kw)
    # the worker's raw sums, arguments as spatialcor builds them (amep/spatialcor.py:262-320)
    bl = box[:, 1] - box[:, 0]
    rmax = kw.get("rmax")
    if rmax is None:
        rmax = max(bl) / 2.
    edges = kw.get("dbin_edges")
    if edges is None:
        edges = np.linspace(0.0, rmax, kw["ndists"]) if kw.get("ndists") is not None else np.arange(0.0, rmax, 1.05)
    pbc = kw.get("pbc", True)
    tree = amep.pbc.kdtree(coords, box, pbc=pbc)
    oc = coords if other is None else other
    ov = values if other_values is None else other_values
    if pbc:
        oc = amep.pbc.fold(oc + bl / 2. - np.mean(box, axis=1), box)
    cor, norm = chunk(0, len(oc), coords, oc, bl, values, ov, edges, tree)
    out = dict(coords=coords, box=box, values=values, c=c, r=r, cor=np.asarray(cor), norm=norm.astype(np.int64),
               edges=np.asarray(edges), pbc=np.bool_(pbc),
               rmax=np.float64(np.nan if kw.get("rmax") is None else kw["rmax"]),
               ndists=np.int64(-1 if kw.get("ndists") is None else kw["ndists"]))
    if other is not None:
        out["other"], out["other_values"] = other, other_values
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: pairs={int(norm.sum())} c.dtype={c.dtype} cor.dtype={np.asarray(cor).dtype} nb={len(c)} c[:3]={c[:3]}")


def main():
    from refshim import load_reference
    amep = load_reference()
    rng = np.random.default_rng(701)

    def flat(n, lo, hi, dtype):
        c = np.zeros((n, 3), dtype=dtype)
        c[:, :2] = rng.uniform(lo, hi, (n, 2)).astype(dtype)
        return c

    # ---- pcf_angle ----------------------------------------------------------------------
    box32 = np.array([[0, 14], [0, 14], [-0.5, 0.5]], dtype=np.float32)
    pcfa_case(amep, "pcfa_f32_same_default", flat(300, 0, 14, np.float32), box32, ndbins=40, nabins=24)
    box64 = np.array([[-6, 8], [2, 13], [-0.5, 0.5]], dtype=np.float64)
    a, b = flat(260, -6, 8, np.float64), flat(150, -6, 8, np.float64)
    a[:, 1] = rng.uniform(2, 13, 260)
    b[:, 1] = rng.uniform(2, 13, 150)
    pcfa_case(amep, "pcfa_f64_cross_psi", a, box64, other=b, ndbins=30, nabins=36, rmax=4.0, psi=np.array([0.3, -0.8]))
    c = flat(220, 0, 14, np.float32)
    ang = rng.uniform(0, 2 * np.pi, 220)
    e = np.stack([np.cos(ang), np.sin(ang), np.zeros(220)], axis=1)
    pcfa_case(amep, "pcfa_f32_orientations", c, box32, ndbins=25, nabins=20, rmax=5.0, e=e)
    pcfa_case(amep, "pcfa_f32_nopbc_psi", flat(250, 0, 14, np.float32), box32, ndbins=32, nabins=16, rmax=6.0, pbc=False,
              psi=np.array([-0.5, 0.2]))
    pcfa_case(amep, "pcfa_f32_box64", flat(200, 0, 14, np.float32), box32.astype(np.float64), ndbins=20, nabins=12, rmax=6.5)

    # ---- spatialcor -------------------------------------------------------------------------
    c = flat(400, 0, 14, np.float32)
    c[:, 2] = 0.0
    scor_case(amep, "scor_f32_scalar_default", c, box32, rng.normal(size=400).astype(np.float32))
    a = rng.uniform(0, 9, (350, 3))
    b = rng.uniform(0, 9, (200, 3))
    box3 = np.array([[0, 9], [0, 9], [0, 9]], dtype=np.float64)
    scor_case(amep, "scor_f64_cross_vector", a, box3, rng.normal(size=(350, 3)), other=b,
              other_values=rng.normal(size=(200, 3)), ndists=25, rmax=4.0)
    c = flat(380, -7, 7, np.float32)
    boxc = np.array([[-7, 7], [-7, 7], [-0.5, 0.5]], dtype=np.float32)
    psi = amep.order.psi_k(c, boxc, k=6)
    scor_case(amep, "scor_f32_complex_psi6", c, boxc, psi, rmax=6.0)
    scor_case(amep, "scor_f32_nopbc_vector", c, boxc, rng.normal(size=(380, 2)).astype(np.float32), pbc=False,
              dbin_edges=np.array([0.0, 0.5, 1.0, 1.7, 2.5, 4.0, 6.0]))


if __name__ == "__main__":
    main()
