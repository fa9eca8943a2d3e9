"""Host-buffer (e2e) path probe: time per frame for device-resident vs host-staged inputs."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amep_b200 import _core, _lib
n, nf = 1_000_000, 128
L = float(np.sqrt(n / 0.6366))
g = torch.Generator(device="cuda").manual_seed(0)
c = torch.zeros((nf, n, 3), device="cuda", dtype=torch.float32)
c[:, :, :2] = (torch.rand((nf, n, 2), generator=g, device="cuda", dtype=torch.float64) * L).float()
box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
edges = np.linspace(0.0, 10.0, 501)
host = torch.empty((nf, n, 3), dtype=torch.float32).pin_memory(); host.copy_(c); hn = host.numpy()
out_h = torch.empty((nf, 500), dtype=torch.int64).pin_memory(); on = out_h.numpy()
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3 / nf
ctx = _lib.context(0); ctx.set_timing(True)
print("device ms/frame", round(t(lambda: _core.rdf_histograms(c, box, edges)), 4), flush=True)
i = ctx.rdf_info(); print("  device-mode info: pair %.4f build %.4f ms/frame, sub_batches %d" % (i["pair_kernel_ms"]/nf, i["build_ms"]/nf, i["sub_batches"]), flush=True)
print("host   ms/frame", round(t(lambda: _core.rdf_histograms(hn, box, edges, out=on)), 4), "stage_mb", os.environ.get("AMEPB_STAGE_MB"), "null_stream", os.environ.get("AMEPB_HOST_NULL_STREAM"), flush=True)
i = ctx.rdf_info(); print("  host-mode info: pair %.4f build %.4f ms/frame, sub_batches %d" % (i["pair_kernel_ms"]/nf, i["build_ms"]/nf, i["sub_batches"]), flush=True)
