"""Golden vectors on the BASELINE configs, produced by the UNMODIFIED reference (AMEP 1.3.0).

Run in the build container only (needs /root/reference):

    python oracle/make_golden_configs.py

``rdf_cfg1_frame0_f32.npz``: frame 0 of BASELINE.json configs[0] (2D uniform, 4096 particles,
float32 trajectory regime, nbins=500, default rmax = L/2 -> float32 edges, the all-images regime):
``amep.spatialcor.rdf`` output plus the integer counts of its chunk worker.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from make_golden import _rdf_case
    from refshim import load_reference
    from baseline_cases import box2d, cfg1_frame
    amep = load_reference()
    c, L = cfg1_frame(0)
    _rdf_case(amep, "rdf_cfg1_frame0_f32", c.astype(np.float32), box2d(L), nbins=500)


if __name__ == "__main__":
    main()
