"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares; host logic (frame selection, sharding, thresholds, planning)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

import amep_oracle as orc


def _lib_path():
    from amep_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.LIB_PATH


def test_library_exports_every_declared_symbol():
    from amep_b200 import _lib
    header = open(os.path.join(ROOT, "include", "amepb.h")).read()
    declared = set(re.findall(r"\b(amepb_[a-z0-9_]+)\s*\(", header))
    declared -= {"amepb_ctx", "amepb_rdf_info"}
    assert declared == set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib_path())
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert _lib.load().amepb_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import amep_b200
    with pytest.raises(amep_b200.AmepbError):
        amep_b200.spatialcor.rdf(np.random.rand(10, 3), np.array([[0, 1], [0, 1], [0, 1.0]]))
    with pytest.raises(amep_b200.AmepbError):
        amep_b200.spatialcor.pcf2d(np.random.rand(10, 3), np.array([[0, 1], [0, 1], [0, 1.0]]), rmax=0.4, nxbins=8, nybins=8)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "amep_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "amep_oracle" not in text and "c_oracle" not in text and "rdf_oracle" not in text, fn


def test_frame_selection_matches_average_func():
    from amep_b200.utils import evaluated_indices
    for n, skip, nav in [(100, 0.0, 100), (51, 0.9, 2), (10, 0.5, 50), (10, 0.0, None), (1000, 0.25, 17), (7, 0.3, 4)]:
        np.testing.assert_array_equal(evaluated_indices(n, skip, nav), orc.frame_indices(n, skip, nav))


def test_shard_bounds_cover_everything():
    from amep_b200.dist import shard_bounds
    for n in (0, 1, 7, 64, 1000):
        for ws in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
                assert b == c and b - a >= d - c >= 0 and (b - a) - (d - c) <= 1


def test_sfiso_direction_dedup():
    from amep_b200.spatialcor import _sfiso_directions, _unique_directions
    u2 = _sfiso_directions(True, 8)
    reps, w = _unique_directions(u2)
    assert len(reps) == 4 and w.sum() == 8
    u3 = _sfiso_directions(False, 8)
    np.testing.assert_array_equal(u3, orc.sfiso_directions(False, 8))
    reps, w = _unique_directions(u3)
    assert len(reps) == 13 and w.sum() == 32


def test_plan_and_thresholds_host_logic():
    """rdf_plan.h compiled on the CPU (tests/host_plan_check.cpp): thresholds
    bin exactly like np.histogram, and the planned grid covers every in-range pair."""
    exe = os.path.join(ROOT, "tests", "_build", "host_plan_check")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    src = os.path.join(ROOT, "tests", "host_plan_check.cpp")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "amep_b200", "csrc"),
                    src, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "ALL OK" in out, out


def test_sq_from_sf2d_matches_reference():
    from amep_b200.utils import sq_from_sf2d
    from conftest import GOLDEN
    z = np.load(os.path.join(GOLDEN, "sq_from_sf2d.npz"))
    sq, q = sq_from_sf2d(z["S"], z["Qx"], z["Qy"])
    np.testing.assert_array_equal(q, z["q"])
    np.testing.assert_array_equal(sq, z["sq"])


def test_tensor_trajectory_protocol():
    """The stand-in trajectory offers what the evaluate classes use of amep's (amep/base.py:429-1080, 1263-1312)."""
    import amep_b200
    rng = np.random.default_rng(3)
    coords = rng.uniform(0, 5, (6, 40, 3)).astype(np.float32)
    types = np.array([1] * 25 + [2] * 15)
    box = np.array([[0, 5], [0, 5], [0, 5]], dtype=np.float32)
    traj = amep_b200.TensorTrajectory(coords, box, times=np.arange(6) * 0.1, types=types)
    assert len(traj) == 6 and traj[2].n() == 40 and traj[2].n(ptype=2) == 15
    np.testing.assert_array_equal(traj[-1].coords(ptype=1), coords[5, :25])
    np.testing.assert_array_equal(traj[1].center, np.mean(box, axis=1))
    frames = traj[np.array([0, 3])]
    assert [f.time for f in frames] == [0.0, 0.30000000000000004]
    stack, b = traj.stack([1, 2, 3], ptype=2)
    np.testing.assert_array_equal(stack, coords[1:4, 25:])
    stack, b = traj.stack([4, 0])
    np.testing.assert_array_equal(stack, coords[[4, 0]])
    with pytest.raises(IndexError):
        traj[6]
