"""torchrun --nproc-per-node N tools/dist_check.py: evaluate.* sharded over N GPUs (NCCL) must equal
the single-rank result bit for bit (RDF) / to rounding (S(q))."""
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
traj = amep_b200.TensorTrajectory(torch.from_numpy(coords).cuda(), box, times=np.arange(nf) * 1.0)
ok = True
for name, make in (("RDF", lambda d: amep_b200.evaluate.RDF(traj, nav=nf, rmax=8.0, nbins=200, distributed=d)),
                   ("SFiso", lambda d: amep_b200.evaluate.SFiso(traj, nav=nf, qmax=6.0, twod=True, num=8, distributed=d)),
                   ("SF2d", lambda d: amep_b200.evaluate.SF2d(traj, nav=nf, rotate=False, qmax=3.0, distributed=d)),
                   ("SF2d fft", lambda d: amep_b200.evaluate.SF2d(traj, nav=nf, rotate=False, qmax=3.0, mode="fft", accuracy=0.2, distributed=d))):
    sharded, alone = make(True), make(False)
    a, b = np.asarray(sharded.frames), np.asarray(alone.frames)
    same = np.array_equal(a, b)
    close = np.allclose(a, b, rtol=1e-12, atol=1e-12 * np.abs(b).max())
    ok = ok and (same if name == "RDF" else close)
    if rank == 0:
        print(f"{name}: world {world}, frames {a.shape}, bit-identical {same}, close {close}", flush=True)
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check", "OK" if int(t.item()) == 1 else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) == 1 else 1)
