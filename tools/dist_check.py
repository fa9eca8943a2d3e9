"""torchrun --nproc-per-node N tools/dist_check.py: the evaluate classes sharded over N GPUs (NCCL)
against the single-rank result -- bit for bit for RDF (integer counts) and for every per-frame row,
to rounding for averages that are reduced as partial sums."""
import os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amep_b200
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(5)
nf, n, L = 11, 3000, 40.0
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
coords = np.zeros((nf, n, 3), dtype=np.float32)
coords[:, :, :2] = rng.uniform(0, L, (nf, n, 2))
vel = rng.normal(size=(nf, n, 3)).astype(np.float32)
traj = amep_b200.TensorTrajectory(torch.from_numpy(coords).cuda(), box, times=np.arange(nf) * 1.0,
                                  velocities=torch.from_numpy(vel).cuda())
host_traj = amep_b200.TensorTrajectory(coords, box, times=np.arange(nf) * 1.0)   # host arrays: every rank stages on its own GPU
ok = True
ev = amep_b200.evaluate


def report(name, same, close):
    global ok
    ok = ok and close
    if rank == 0:
        print(f"{name}: world {world}, bit-identical {same}, close {close}", flush=True)


cases = (
    ("RDF", lambda d: ev.RDF(traj, nav=nf, rmax=8.0, nbins=200, distributed=d), True),
    ("RDF host arrays", lambda d: ev.RDF(host_traj, nav=nf, rmax=8.0, nbins=200, distributed=d), True),
    ("RDF one frame (sharded inside the frame)", lambda d: ev.RDF(traj, nav=1, skip=0.5, rmax=8.0, nbins=200, distributed=d), True),
    ("SFiso", lambda d: ev.SFiso(traj, nav=nf, qmax=6.0, twod=True, num=8, distributed=d), False),
    ("SFiso one frame (particles split over the ranks)", lambda d: ev.SFiso(traj, nav=1, qmax=6.0, twod=True, num=8, distributed=d), False),
    ("SF2d", lambda d: ev.SF2d(traj, nav=nf, rotate=False, qmax=3.0, distributed=d), False),
    ("SF2d rotate=True", lambda d: ev.SF2d(traj, nav=5, qmax=3.0, distributed=d), False),
    ("SF2d one frame (rows of q tiles sharded over the ranks)", lambda d: ev.SF2d(traj, nav=1, rotate=False, qmax=6.0, distributed=d), False),
    ("SF2d rotate=True, one frame (sharded over q)", lambda d: ev.SF2d(traj, nav=1, qmax=6.0, distributed=d), False),
    ("SF2d fft", lambda d: ev.SF2d(traj, nav=nf, rotate=False, qmax=3.0, mode="fft", accuracy=0.2, distributed=d), False),
    ("PCF2d (psi_6 orientation)", lambda d: ev.PCF2d(traj, nav=4, nxbins=50, nybins=50, rmax=6.0, distributed=d), False),
    ("PCFangle one frame (query particles sharded over the ranks)", lambda d: ev.PCFangle(traj, nav=1, mode="psi6", ndbins=60, nabins=24, rmax=8.0, distributed=d), False),
    ("SpatialVelCor", lambda d: ev.SpatialVelCor(traj, nav=4, distributed=d), False),
    ("SpatialVelCor one frame (query particles sharded over the ranks)", lambda d: ev.SpatialVelCor(traj, nav=1, distributed=d), False),
    ("HexOrderCor one frame (query particles sharded over the ranks)", lambda d: ev.HexOrderCor(traj, nav=1, distributed=d), False),
    ("PCF2d one frame (query particles sharded over the ranks)", lambda d: ev.PCF2d(traj, nav=1, nxbins=50, nybins=50, rmax=6.0, distributed=d), False),
)
for name, make, exact in cases:
    sharded, alone = make(True), make(False)
    a, b = np.asarray(sharded.frames), np.asarray(alone.frames)
    same = np.array_equal(a, b)
    avg_close = np.allclose(sharded.avg, alone.avg, rtol=1e-12, atol=1e-12 * np.abs(alone.avg).max())
    if exact:
        same = same and np.array_equal(sharded.avg, alone.avg) and np.array_equal(sharded.counts, alone.counts)
    report(name, same, (same if exact else np.allclose(a, b, rtol=1e-12, atol=1e-12 * np.abs(b).max())) and avg_close)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check", "OK" if int(t.item()) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
