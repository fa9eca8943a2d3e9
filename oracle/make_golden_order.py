"""Golden vectors of the psi_6 orientation prologue, produced by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_order.py

For every case: ``amep.order.k_nearest_neighbors`` ids and ``amep.order.psi_k`` (k = 6) of the frame,
then the body of ``amep.evaluate.SF2d.__compute_particles`` with ``rotate=True``
(amep/evaluate.py:1349-1394) executed step by step with the reference's own functions: mean psi_6,
the orientation angle, ``rotate_coords(pbc_points(...), -theta, center)``, ``in_box`` and ``sf2d``.
``amep.evaluate.PCF2d.__compute`` (amep/evaluate.py:711-745) is covered by the same psi.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def hex_lattice(nx, ny, a, jitter, rng, rows_along_y=False):
    """Triangular lattice of nx x ny (ny even) sites, periodic in a box nx*a by ny*a*sqrt(3)/2."""
    ix, iy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    x = (ix + 0.5 * (iy % 2)) * a
    y = iy * a * np.sqrt(3) / 2
    pts = np.stack([x.ravel(), y.ravel()], axis=1)
    lx, ly = nx * a, ny * a * np.sqrt(3) / 2
    pts += rng.normal(0, jitter * a, pts.shape)
    pts[:, 0] = np.mod(pts[:, 0], lx)
    pts[:, 1] = np.mod(pts[:, 1], ly)
    if rows_along_y:
        pts = pts[:, ::-1].copy()
        lx, ly = ly, lx
    out = np.zeros((len(pts), 3))
    out[:, :2] = pts
    return out, lx, ly


def case(amep, name, coords, box, qmax, other=None):
    from amep.order import k_nearest_neighbors, psi_k
    from amep.pbc import pbc_points
    from amep.utils import in_box, rotate_coords
    _, _, ids = k_nearest_neighbors(coords, box, k=6, pbc=True, enforce_nd=2)
    ids = np.array([list(r) for r in ids]).astype(int)
    psi_all = psi_k(coords, box, k=6)
    # amep/evaluate.py:1352-1363
    psi = np.mean(psi_all)
    psi2 = np.array([psi.real, psi.imag])
    ex = np.array([1, 0])
    theta = np.arccos(np.dot(ex, psi2) / np.sqrt(psi2[0] ** 2 + psi2[1] ** 2))
    if psi2[1] < 0:
        theta = 2 * np.pi - theta
    center = np.mean(box, axis=1)
    # amep/evaluate.py:1366-1377
    rc = in_box(rotate_coords(pbc_points(coords, box), -theta, center), box)
    ro = rc if other is None else in_box(rotate_coords(pbc_points(other, box), -theta, center), box)
    S, qx, qy = amep.spatialcor.sf2d(rc, box, other_coords=ro, qmax=qmax)
    out = dict(coords=coords, box=box, neighbors=ids, psi=psi_all, psi_mean=np.array([psi]), theta=np.float64(theta),
               rotated=rc, S=S, Qx=qx, Qy=qy, qmax=np.float64(qmax))
    if other is not None:
        out["other"] = other
        out["rotated_other"] = ro
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: N={len(coords)} psi.dtype={psi_all.dtype} mean={psi} theta={theta} kept={len(rc)} "
          f"S.dtype={S.dtype} grid={S.shape} rotated.dtype={rc.dtype}")


def main():
    from refshim import load_reference
    amep = load_reference()
    rng = np.random.default_rng(601)
    c, lx, ly = hex_lattice(24, 28, 1.1, 0.06, rng)
    box = np.array([[0, lx], [0, ly], [-0.5, 0.5]])
    case(amep, "order_hex_f32", c.astype(np.float32), box.astype(np.float32), qmax=3.0)
    c, lx, ly = hex_lattice(20, 24, 1.0, 0.12, rng, rows_along_y=True)
    box = np.array([[0, lx], [0, ly], [-0.5, 0.5]])
    case(amep, "order_hex_rows_y_f64", c, box, qmax=3.5)
    # uniform random points (|psi_6| ~ 1/sqrt(N): the angle is arbitrary), some outside the box,
    # a box that does not start at the origin, a second species for the cross structure factor
    c = np.zeros((700, 3))
    c[:, 0] = rng.uniform(-11.0, 12.5, 700)
    c[:, 1] = rng.uniform(3.5, 20.0, 700)
    box = np.array([[-10, 12], [4, 19], [-0.5, 0.5]])
    o = np.zeros((300, 3))
    o[:, 0] = rng.uniform(-10, 12, 300)
    o[:, 1] = rng.uniform(4, 19, 300)
    case(amep, "order_uniform_spill_f32", c.astype(np.float32), box.astype(np.float32), qmax=2.5, other=o.astype(np.float32))
    case(amep, "order_uniform_spill_f64", c, box, qmax=2.5)


if __name__ == "__main__":
    main()
