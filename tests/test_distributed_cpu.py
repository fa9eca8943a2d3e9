"""world_size-2 `gloo` tests of the frame sharding + all-reduce host logic (no GPU).

The per-frame kernels are replaced by the CPU oracle (test infrastructure) so that
what is exercised is the product's own sharding, table assembly and averaging code."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def orc_pcf2d(coords, box, psi):
    import amep_oracle as orc
    return orc.pcf2d(coords, box, None, 12, 12, 3.0, psi)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _oracle_rdf_frames(coords, box, other_coords=None, nbins=None, rmax=None, mode="diff", pbc=True,
                       return_counts=False, reduce=None):
    import amep_oracle as orc
    coords = np.asarray(coords)
    g_rows, r_rows, h_rows = [], [], []
    for f in range(coords.shape[0]):
        b = box[f] if np.ndim(box) == 3 else box
        o = None if other_coords is None else np.asarray(other_coords)[f]
        hist, edges = orc.rdf_counts(coords[f], b, o, nbins=nbins, rmax=rmax, pbc=pbc)
        g, r = orc.rdf_normalise(hist, edges, coords[f], b, len(coords[f]) if o is None else len(o))
        g_rows.append(g), r_rows.append(r), h_rows.append(hist)
    out = (np.array(g_rows), np.array(r_rows))
    return out + (np.array(h_rows),) if return_counts else out


def _oracle_pcf2d_frames(coords, box, other_coords=None, nxbins=None, nybins=None, rmax=None, psi=None,
                         zero_last_column=True):
    import amep_oracle as orc
    coords = np.asarray(coords)
    out = []
    for f in range(coords.shape[0]):
        b = box[f] if np.ndim(box) == 3 else box
        o = None if other_coords is None else np.asarray(other_coords)[f]
        out.append(orc.pcf2d(coords[f], b, o, nxbins, nybins, rmax, None if psi is None else psi[f]))
    return tuple(np.array([r[k] for r in out]) for k in range(3))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import amep_b200
    from amep_b200 import dist as adist, evaluate

    # (1) assemble_rows: integer rows and complex rows
    lo, hi = adist.shard_bounds(7, rank, world)
    rows = np.arange(7 * 5, dtype=np.int64).reshape(7, 5) * 3 + 1
    table = adist.assemble_rows(rows[lo:hi], lo, 7)
    assert np.array_equal(table, rows)
    crow = (np.arange(7 * 4).reshape(7, 4) * (1 + 2j)).astype(np.complex64)
    ctab = adist.assemble_rows(crow[lo:hi], lo, 7)
    assert np.array_equal(ctab, crow)

    # (2) evaluate.RDF over a sharded trajectory equals the reference's average_func result
    evaluate.rdf_frames = _oracle_rdf_frames
    z = np.load(os.path.join(GOLDEN, "evaluate_rdf_f32.npz"))
    traj = amep_b200.TensorTrajectory(z["coords"], z["box"])
    ev = evaluate.RDF(traj, skip=float(z["skip"]), nav=int(z["nav"]), nbins=int(z["nbins"]))
    assert np.array_equal(ev.indices, z["indices"])
    assert np.array_equal(ev.frames, z["frames"])
    assert np.array_equal(ev.avg, z["avg"][0]) and np.array_equal(ev.r, z["avg"][1])
    # more ranks than work for one of them: nav=1
    ev1 = evaluate.RDF(traj, skip=0.0, nav=1, nbins=50)
    solo, _ = _oracle_rdf_frames(z["coords"][ev1.indices], z["box"], nbins=50)
    assert np.array_equal(ev1.avg, solo[0])
    # (3) evaluate.PCF2d: frames sharded, one all-reduce, average as average_func does
    evaluate.pcf2d_frames = _oracle_pcf2d_frames
    rng = np.random.default_rng(77)
    pc = np.zeros((5, 150, 3), dtype=np.float32)
    pc[:, :, :2] = rng.uniform(-6, 6, (5, 150, 2)).astype(np.float32)
    pc[:, :, 2] = 0.25
    pbox = np.array([[-6, 6], [-6, 6], [-0.5, 0.5]], dtype=np.float32)
    psi = rng.normal(size=(5, 2))
    ptraj = amep_b200.TensorTrajectory(pc.copy(), pbox)
    evp = evaluate.PCF2d(ptraj, skip=0.0, nav=4, psi=psi, nxbins=12, nybins=12, rmax=3.0)
    want = [orc_pcf2d(pc[i], pbox, psi[i]) for i in evp.indices]
    assert np.array_equal(evp.avg, np.mean(np.array([w[0] for w in want]), axis=0))
    assert np.array_equal(evp.x, want[0][1]) and np.array_equal(evp.y, want[0][2])
    assert evp.frames.shape == (len(evp.indices), 3, 12, 12)
    np.save(os.path.join(tmpdir, f"ok{rank}.npy"), np.array([1]))
    dist.destroy_process_group()


def test_gloo_world2_sharding_and_allreduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0.npy") and os.path.exists(tmp_path / "ok1.npy")
