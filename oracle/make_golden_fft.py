"""Golden vectors for sf2d / sfiso with mode='fft', produced by the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_fft.py

Stores the seeded inputs and what ``amep.spatialcor.sf2d(mode='fft')`` /
``amep.spatialcor.sfiso(mode='fft')`` returned (amep/spatialcor.py:1667-1685, 1911-1940), plus
the integer counts of the histogram density estimate (``amep.continuum.coords_to_density``,
amep/continuum.py:51-139, 300-407) that the FFT starts from.
"""
import glob
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def _case(amep, name, coords, box, qmax, accuracy, other=None, ntot=None):
    S, Qx, Qy = amep.spatialcor.sf2d(coords, box, other_coords=other, qmax=qmax, mode="fft", accuracy=accuracy, Ntot=ntot)
    dmin = 2 * np.pi / 2 / qmax / (accuracy * 10)
    d, X, Y = amep.continuum.coords_to_density(coords, box, dmin=dmin)
    counts = np.rint(d * dmin ** 2).astype(np.int64)
    out = dict(coords=coords, box=box, S=S, Qx=Qx, Qy=Qy, counts=counts, qmax=np.float64(qmax),
               accuracy=np.float64(accuracy), ntot=np.int64(-1 if ntot is None else ntot))
    if other is not None:
        out["other"] = other
    s_iso, q_iso = amep.spatialcor.sfiso(coords, box, len(coords) if ntot is None else ntot, qmax=qmax, twod=True,
                                         mode="fft", accuracy=accuracy, other_coords=other)
    out["S_iso"], out["q_iso"] = s_iso, q_iso
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    print(f"{name}: grid {counts.shape}, sum(counts)={counts.sum()} of {len(coords)}, S {S.shape} {S.dtype}, "
          f"max S {S.max():.4g}, sfiso nq={len(q_iso)}")


def main():
    from refshim import load_reference
    amep = load_reference()
    rng = np.random.default_rng(301)
    box = np.array([[0, 30], [0, 30], [-0.5, 0.5]], dtype=np.float64)
    c = np.zeros((1500, 3))
    c[:, :2] = rng.uniform(0, 30, (1500, 2))
    c[0, :2] = (0.0, 0.0)          # on the lower box corner
    c[1, :2] = (30.0, 30.0)        # on the upper box corner
    c[2, :2] = (30.0, 7.5)
    o = np.zeros((900, 3))
    o[:, :2] = rng.uniform(0, 30, (900, 2))
    _case(amep, "sf2d_fft_f32", c.astype(np.float32), box.astype(np.float32), qmax=6.0, accuracy=0.2)
    _case(amep, "sf2d_fft_f64_cross", c, box, qmax=5.0, accuracy=0.1, other=o, ntot=2400)
    boxr = np.array([[-10, 25], [-10, 25], [-0.5, 0.5]], dtype=np.float32)
    cr = np.zeros((1200, 3), dtype=np.float32)
    cr[:, :2] = rng.uniform(-10, 25, (1200, 2)).astype(np.float32)
    cr[5, 0] = 26.0                # outside the box: dropped or in the margin bin, as np.histogram2d decides
    cr[6, 1] = -10.4
    _case(amep, "sf2d_fft_f32_offset", cr, boxr, qmax=4.0, accuracy=0.3)
    dumps = sorted(glob.glob(os.path.join(amep.__path__[0], "..", "examples", "data", "lammps", "dump*.txt")),
                   key=lambda p: int(os.path.basename(p)[4:-4]))
    from make_golden import _read_lammps
    cl, bl = _read_lammps(dumps[-1])
    _case(amep, "sf2d_fft_lammps_last_f32", cl.astype(np.float32), bl.astype(np.float32), qmax=8.0, accuracy=0.1)


if __name__ == "__main__":
    main()
