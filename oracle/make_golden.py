"""Generate tests/golden/*.npz by running the UNMODIFIED reference (AMEP 1.3.0).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Each fixture stores the seeded inputs and what the reference returned for them:
``amep.spatialcor.rdf`` (g, r) plus the integer counts of its chunk worker
``__rdf_diff`` (amep/spatialcor.py:348-388) on the images of ``pbc_points``;
``amep.spatialcor.sfiso(mode='fast')``; ``amep.spatialcor.sf2d(mode='std')``;
and an ``evaluate.RDF``-equivalent run of the reference's own
``amep.utils.average_func`` over in-memory frames (SURVEY.md section 8c -- real
``ParticleTrajectory`` objects need h5py, which this image lacks).

The fixtures travel to the GPU box; /root/reference does not.
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def _rdf_case(amep, name, coords, box, other=None, **kw):
    """Reference g(r), r and integer counts for one frame."""
    rdf_diff = getattr(amep.spatialcor, "__rdf_diff")
    pbc = kw.get("pbc", True)
    g, r = amep.spatialcor.rdf(coords, box, other_coords=other, mode="diff", **kw)
    queries = coords if other is None else other
    nbins = kw.get("nbins") or 500
    rmax = kw.get("rmax")
    if rmax is None:
        rmax = max(box[:, 1] - box[:, 0]) / 2
    edges = np.linspace(0.0, rmax, nbins + 1)
    targets = amep.pbc.pbc_points(coords, box, width=0.5, fold_coords=False) if pbc else coords
    counts = rdf_diff(0, len(queries), targets, queries, box, edges, pbc)
    out = dict(coords=coords, box=box, g=g, r=r, counts=counts.astype(np.int64),
               edges=edges, n_images=np.int64(len(targets)),
               nbins=np.int64(-1 if kw.get("nbins") is None else kw["nbins"]),
               rmax=np.float64(np.nan if kw.get("rmax") is None else kw["rmax"]),
               rmax_is_default=np.bool_(kw.get("rmax") is None),
               pbc=np.bool_(pbc))
    if other is not None:
        out["other"] = other
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: sum(counts)={counts.sum()} images={len(targets)} g.dtype={g.dtype} r.dtype={r.dtype}")


def _read_lammps(path):
    with open(path) as fh:
        lines = fh.read().split("\n")
    n = int(lines[3])
    box = np.array([[float(x) for x in lines[5 + i].split()] for i in range(3)])
    cols = lines[8].split()[2:]
    ix, iy, iz = cols.index("x"), cols.index("y"), cols.index("z")
    data = np.array([[float(v) for v in ln.split()] for ln in lines[9:9 + n]])
    order = np.argsort(data[:, 0])
    data = data[order]
    return data[:, [ix, iy, iz]], box


def main():
    from refshim import load_reference
    amep = load_reference()
    os.makedirs(GOLDEN, exist_ok=True)
    versions = dict(amep=amep.__version__, numpy=np.__version__)
    import scipy, numba
    versions.update(scipy=scipy.__version__, numba=numba.__version__)
    with open(os.path.join(GOLDEN, "VERSIONS.txt"), "w") as fh:
        for k, v in versions.items():
            fh.write(f"{k} {v}\n")

    # ---- RDF ---------------------------------------------------------------
    # (1) the reference's own unit-test input (test/test_spatialcor.py:52-102)
    rng = np.random.default_rng(0)
    c = np.zeros((100, 3))
    c[:, :2] = rng.uniform(low=-10, high=10, size=(100, 2))
    box = np.array([[-10, 10], [-10, 10], [-0.5, 0.5]])
    _rdf_case(amep, "rdf_unittest_2d_f64", c, box, nbins=50, rmax=5.0)
    gk, _ = amep.spatialcor.rdf(c, box, nbins=50, pbc=True, rmax=5.0, mode="kdtree")
    np.savez_compressed(os.path.join(GOLDEN, "rdf_unittest_2d_kdtree.npz"), g=gk)

    # (2,3) non-square 2D box, particles outside the box, default rmax > min(L)/2
    rng = np.random.default_rng(101)
    c = np.zeros((700, 3))
    c[:, 0] = rng.uniform(-2.0, 33.0, 700)     # box x: [0,30]
    c[:, 1] = rng.uniform(-1.5, 21.0, 700)     # box y: [0,20]
    box = np.array([[0, 30], [0, 20], [-0.5, 0.5]], dtype=np.float64)
    _rdf_case(amep, "rdf_2d_nonsquare_f64", c, box)
    _rdf_case(amep, "rdf_2d_nonsquare_f32", c.astype(np.float32), box.astype(np.float32))

    # (4) trajectory regime but a Python-float rmax (float64 edges, float32 arithmetic)
    rng = np.random.default_rng(102)
    c = np.zeros((900, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, 25, (900, 2)).astype(np.float32)
    box = np.array([[0, 25], [0, 25], [-0.5, 0.5]], dtype=np.float32)
    _rdf_case(amep, "rdf_2d_f32_rmax_py", c, box, rmax=4.0, nbins=200)

    # (5) 2D with a non-zero constant z given in float64 (dz = z - fl32(z) != 0)
    c64 = c.astype(np.float64)
    c64[:, 2] = 0.1
    _rdf_case(amep, "rdf_2d_f64_zconst", c64, box.astype(np.float64), rmax=3.0, nbins=64)

    # (6) 3D cross-species, explicit rmax / nbins
    rng = np.random.default_rng(103)
    a = rng.uniform(0, 12, (600, 3))
    b = rng.uniform(0, 12, (350, 3))
    box = np.array([[0, 12], [0, 12], [0, 12]], dtype=np.float64)
    _rdf_case(amep, "rdf_3d_cross_f64", a, box, other=b, rmax=5.5, nbins=137)
    _rdf_case(amep, "rdf_3d_cross_f32", a.astype(np.float32), box.astype(np.float32),
              other=b.astype(np.float32), rmax=5.5, nbins=137)

    # (7) 3D default rmax (= L/2), non-cubic box with an offset origin
    rng = np.random.default_rng(104)
    box = np.array([[-3, 7], [2, 10], [-4, 5]], dtype=np.float64)
    a = rng.uniform(box[:, 0], box[:, 1], (500, 3))
    _rdf_case(amep, "rdf_3d_default_f64", a, box)
    _rdf_case(amep, "rdf_3d_default_f32", a.astype(np.float32), box.astype(np.float32))

    # (8) no periodic boundary conditions
    _rdf_case(amep, "rdf_3d_nopbc_f64", a, box, pbc=False, rmax=4.0, nbins=80)
    _rdf_case(amep, "rdf_3d_nopbc_f32", a.astype(np.float32), box.astype(np.float32),
              pbc=False, rmax=4.0, nbins=80)

    # (9) real data: two frames of the reference's LAMMPS ABP example
    #     (examples/data/lammps/dump*.txt, 4000 particles, box 79.266^2), stored
    #     as float32 like an h5amep trajectory (amep/base.py:126)
    dumps = sorted(glob.glob(os.path.join(amep.__path__[0], "..", "examples", "data",
                                          "lammps", "dump*.txt")),
                   key=lambda p: int(os.path.basename(p)[4:-4]))
    for tag, path in (("first", dumps[0]), ("last", dumps[-1])):
        c, box = _read_lammps(path)
        _rdf_case(amep, f"rdf_lammps_{tag}_f32", c.astype(np.float32),
                  box.astype(np.float32), rmax=10.0, nbins=500)

    # ---- sfiso(mode='fast') ---------------------------------------------------
    rng = np.random.default_rng(201)
    box2 = np.array([[0, 18], [0, 18], [-0.5, 0.5]], dtype=np.float64)
    c2 = np.zeros((500, 3))
    c2[:, :2] = rng.uniform(0, 18, (500, 2))
    for tag, cc, bb in (("f64", c2, box2), ("f32", c2.astype(np.float32), box2.astype(np.float32))):
        s, q = amep.spatialcor.sfiso(cc, bb, len(cc), qmax=12.0, twod=True, mode="fast", num=8)
        np.savez_compressed(os.path.join(GOLDEN, f"sfiso_2d_{tag}.npz"), coords=cc, box=bb,
                            S=s, q=q, qmax=12.0, twod=True, num=8)
        print(f"sfiso_2d_{tag}: nq={len(q)} S[:3]={s[:3]}")
    rng = np.random.default_rng(202)
    box3 = np.array([[0, 9], [0, 9], [0, 9]], dtype=np.float64)
    a = rng.uniform(0, 9, (450, 3))
    b = rng.uniform(0, 9, (300, 3))
    s, q = amep.spatialcor.sfiso(a, box3, len(a), qmax=15.0, twod=False, mode="fast", num=8)
    np.savez_compressed(os.path.join(GOLDEN, "sfiso_3d_f64.npz"), coords=a, box=box3,
                        S=s, q=q, qmax=15.0, twod=False, num=8)
    s, q = amep.spatialcor.sfiso(a.astype(np.float32), box3.astype(np.float32), len(a), qmax=15.0,
                                 twod=False, mode="fast", num=8, other_coords=b.astype(np.float32))
    np.savez_compressed(os.path.join(GOLDEN, "sfiso_3d_cross_f32.npz"), coords=a.astype(np.float32),
                        other=b.astype(np.float32), box=box3.astype(np.float32),
                        S=s, q=q, qmax=15.0, twod=False, num=8)
    print(f"sfiso_3d: nq={len(q)}")

    # ---- sfiso(mode='std'), Debye formula ------------------------------------------
    rng = np.random.default_rng(211)
    c2s = np.zeros((300, 3))
    c2s[:, :2] = rng.uniform(0, 18, (300, 2))
    c2s[7] = c2s[3]                                   # a coincident pair (zero distance)
    s, q = amep.spatialcor.sfiso(c2s, box2, len(c2s), qmax=10.0, twod=True, mode="std")
    np.savez_compressed(os.path.join(GOLDEN, "sfiso_std_2d_f64.npz"), coords=c2s, box=box2, S=s, q=q,
                        qmax=10.0, twod=True)
    a3 = rng.uniform(0, 9, (260, 3)).astype(np.float32)
    b3 = rng.uniform(0, 9, (150, 3)).astype(np.float32)
    s, q = amep.spatialcor.sfiso(a3, box3.astype(np.float32), len(a3), qmax=12.0, twod=False, mode="std",
                                 other_coords=b3, chunksize=64)
    np.savez_compressed(os.path.join(GOLDEN, "sfiso_std_3d_cross_f32.npz"), coords=a3, other=b3,
                        box=box3.astype(np.float32), S=s, q=q, qmax=12.0, twod=False, chunksize=64)
    print(f"sfiso_std: nq={len(q)} dtype={s.dtype}")

    # ---- sf2d(mode='std') -----------------------------------------------------
    rng = np.random.default_rng(301)
    box2 = np.array([[0, 20], [0, 20], [-0.5, 0.5]], dtype=np.float64)
    c2 = np.zeros((400, 3))
    c2[:, :2] = rng.uniform(0, 20, (400, 2))
    o2 = np.zeros((250, 3))
    o2[:, :2] = rng.uniform(0, 20, (250, 2))
    s, qx, qy = amep.spatialcor.sf2d(c2, box2, qmax=3.0)
    np.savez_compressed(os.path.join(GOLDEN, "sf2d_f64_onechunk.npz"), coords=c2, box=box2,
                        S=s, Qx=qx, Qy=qy, qmax=3.0, chunksize=-1)
    s, qx, qy = amep.spatialcor.sf2d(c2.astype(np.float32), box2.astype(np.float32), qmax=3.0,
                                     chunksize=37)
    np.savez_compressed(os.path.join(GOLDEN, "sf2d_f32_chunks.npz"), coords=c2.astype(np.float32),
                        box=box2.astype(np.float32), S=s, Qx=qx, Qy=qy, qmax=3.0, chunksize=37)
    s, qx, qy = amep.spatialcor.sf2d(c2, box2, other_coords=o2, qmax=3.0, chunksize=64)
    np.savez_compressed(os.path.join(GOLDEN, "sf2d_f64_cross.npz"), coords=c2, other=o2, box=box2,
                        S=s, Qx=qx, Qy=qy, qmax=3.0, chunksize=64)
    print(f"sf2d: grid={s.shape} dtype={s.dtype}")
    z = np.load(os.path.join(GOLDEN, "sf2d_f64_onechunk.npz"))
    sq, qr = amep.utils.sq_from_sf2d(z["S"], z["Qx"], z["Qy"])
    np.savez_compressed(os.path.join(GOLDEN, "sq_from_sf2d.npz"), S=z["S"], Qx=z["Qx"], Qy=z["Qy"], sq=sq, q=qr)

    # ---- evaluate.RDF semantics via the reference's average_func -----------------
    class Frame:
        def __init__(self, c, b):
            self._c, self.box = c, b

        def coords(self, ptype=None):
            return self._c

    rng = np.random.default_rng(401)
    boxf = np.array([[0, 16], [0, 16], [-0.5, 0.5]], dtype=np.float32)
    stack = np.zeros((7, 300, 3), dtype=np.float32)
    stack[:, :, :2] = rng.uniform(0, 16, (7, 300, 2)).astype(np.float32)
    frames = np.empty(7, dtype=object)
    for i in range(7):
        frames[i] = Frame(stack[i], boxf)
    res, avg, idx = amep.utils.average_func(
        lambda f: amep.spatialcor.rdf(f.coords(), f.box, nbins=100), frames,
        skip=0.3, nr=4, indices=True)
    np.savez_compressed(os.path.join(GOLDEN, "evaluate_rdf_f32.npz"), coords=stack, box=boxf,
                        frames=res, avg=avg, indices=idx, skip=0.3, nav=4, nbins=100)
    print(f"evaluate_rdf: indices={idx} frames.shape={res.shape}")


if __name__ == "__main__":
    main()
