#!/usr/bin/env python
"""Benchmark of the RDF / S(q) hot path (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A *step* is one pass of the hot path over one batch of synthetic frames.
Headline workload at every N (weak scaling, frames sharded over ranks, one NCCL
all-reduce of the per-frame histograms): BASELINE.json configs[1], "2D active
Brownian particles, 1M particles, RDF rmax=10 sigma, nbins=500" -- uniform
positions at packing fraction 0.5 (rho = 0.6366, L = 1253.3), `--frames` frames
per GPU (default 1000, the config's count; every frame is the same amount of work
repeated).

One JSON line on stdout (rank 0):
  value      in-range ordered pair evaluations per second (sum of all histogram
             counts / time), inputs resident in HBM, all ranks, all-reduce included
  e2e        the same metric through the public API, evaluate.RDF(trajectory over PINNED
             HOST arrays): host->device copies of every frame, kernels, device->host of the
             histograms, the NumPy normalisation and the frame average inside the timed region
  roofline   the pair kernel (dominant launch): algorithmic flops (5 per in-range
             2D pair: 2 SUB, 2 MUL, 1 ADD -- SURVEY.md 8d) over its CUDA-event time,
             against the FP32 FFMA peak measured by tools/microbench (`frac`); the
             non-fused FMUL+FADD rate the bit-exact arithmetic is limited to is `frac_unfused`
  cpu_baseline  the C restatement of the reference algorithm (oracle/rdf_oracle.c,
             brute force over all periodic images like amep/spatialcor.py:348-388)
             on the host cores, on a bounded sample of the same frame
  strong     BASELINE.json configs[4]: 3D uniform 16.7M particles, 64 frames in total split
             over the N ranks (strong scaling), RDF rmax=10, one all-reduce of the histograms
  sq         S(q) particle*q terms per second for sfiso(mode='fast') and sf2d(mode='std'), each
             with the FP64-pipe roofline of its kernel and a CPU baseline (C oracle, host cores)
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RHO = 0.6366            # packing fraction 0.5 of unit discs
RMAX = 10.0
NBINS = 500
FLOPS_PER_PAIR_2D = 5


def box_length(n):
    return float(np.sqrt(n / RHO))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="amep_b200", choices=["amep_b200", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--frames", type=int, default=1000, help="frames per GPU and step")
    ap.add_argument("--e2e-frames", type=int, default=0,
                    help="frames per GPU of the end-to-end leg (0 = all --frames, bounded by host memory)")
    ap.add_argument("--strong-frames", type=int, default=64, help="total frames of the strong-scaling section")
    ap.add_argument("--strong-particles", type=int, default=1 << 24)
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling section (configs[4])")
    ap.add_argument("--cpu-queries", type=int, default=0, help="query sample of the CPU baseline (0 = auto)")
    ap.add_argument("--no-sq", action="store_true", help="skip the S(q) section")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_first_sample(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc is not None and time.perf_counter() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                return
            time.sleep(0.05)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    smax.append(float(p[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                     p[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(smax)), samples=len(sm),
                       reasons=sorted(reasons))
        return out


# ----------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference algorithm on the host cores
# ----------------------------------------------------------------------------------------
def cpu_frame(n, seed=0):
    """One synthetic frame (float32, trajectory regime) identical in law to the GPU frames."""
    rng = np.random.default_rng([2, seed])
    L = box_length(n)
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, L, (n, 2)).astype(np.float32)
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    return c, box


def cpu_reference_rate(n, nqueries, threads, repeats=1):
    """pair-evals/s of the C restatement of rdf(mode='diff') on `nqueries` queries of one
    frame against all of its periodic images (what the reference evaluates per query)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as corc
    c, box = cpu_frame(n)
    images = corc.images(c, box)
    edges = np.ascontiguousarray(np.linspace(0.0, RMAX, NBINS + 1))
    hist = np.zeros(NBINS, dtype=np.int64)
    import ctypes
    best, pairs = None, 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        corc.lib().oracle_rdf_hist(images.ctypes.data, 0, len(images), c.ctypes.data, 0, 0, nqueries,
                                   edges.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), NBINS,
                                   hist.ctypes.data, threads)
        dt = time.perf_counter() - t0
        pairs = int(hist.sum())
        best = dt if best is None else min(best, dt)
    return pairs / best, best, pairs, len(images)


def auto_cpu_queries(n, threads):
    # ~ 2.5e8 distance evaluations per second and thread; aim at ~12 s
    per_query = 4.0 * n
    return int(max(threads * 4, min(n, 12.0 * 2.5e8 * threads / per_query)))


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port, all host threads),
    each step a bounded query sample of one frame of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as corc
    # all host cores of this process (torchrun exports OMP_NUM_THREADS=1; the oracle sets its
    # own thread count explicitly)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = args.particles
    nq = args.cpu_queries or max(threads * 4, int(1.5 * auto_cpu_queries(n, threads)) // max(1, args.steps + args.warmup))
    for _ in range(args.warmup):
        cpu_reference_rate(n, nq, threads)
    times, pairs = [], 0
    for _ in range(args.steps):
        rate, dt, pairs, nimg = cpu_reference_rate(n, nq, threads)
        times.append(dt)
    dt = float(np.sum(times))
    value = pairs * args.steps / dt
    sample = f"{nq} query particles of one frame x all {nimg} periodic images per step (brute force, as the reference)"
    line = {
        "impl": "reference", "metric": "rdf_pair_evals_per_s", "value": value, "unit": "pair-evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "pair-evals/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pair-evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    n = args.particles
    return {"workload": "BASELINE.json configs[1]: 2D active-Brownian-like particles, RDF rmax=10 sigma, nbins=500, "
                        "periodic box, float32 trajectory regime (uniform positions, packing fraction 0.5)",
            "particles": n, "frames_per_gpu": args.frames, "frames_total": args.frames * world,
            "box_length": box_length(n), "rmax": RMAX, "nbins": NBINS, "parallelism": f"frames sharded over {world} GPU(s)",
            "l2": "inputs larger than L2 (12 MB per frame, %d MB per step and GPU)" % (args.frames * n * 12 // 2 ** 20)}


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from amep_b200 import _core, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n, nf = args.particles, args.frames
    L = box_length(n)
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    edges = np.linspace(0.0, RMAX, NBINS + 1)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    coords = torch.zeros((nf, n, 3), device=dev, dtype=torch.float32)
    chunk = max(1, min(nf, (1 << 28) // (n * 2)))
    for f0 in range(0, nf, chunk):
        f1 = min(nf, f0 + chunk)
        coords[f0:f1, :, :2] = (torch.rand((f1 - f0, n, 2), generator=gen, device=dev, dtype=torch.float64) * L).float()
    ctx = _lib.context(local)
    ctx.set_timing(True)
    table = torch.zeros((nf * world, NBINS), dtype=torch.int64, device=dev)
    local_rows = table[rank * nf:(rank + 1) * nf]

    def step():
        if world > 1:
            table.zero_()
        _core.rdf_histograms(coords, box, edges, out=local_rows)
        if world > 1:
            dist.all_reduce(table, op=dist.ReduceOp.SUM)

    # nvidia-smi is started (and has written its first sample) before the warm-up, so that its
    # start-up cost never lands inside the timed region; it then samples warm-up + timed steps.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pair_ms = build_ms = 0.0
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
        info = ctx.rdf_info()          # synchronises; also collects the pair-kernel CUDA-event time
        pair_ms += info["pair_kernel_ms"]
        build_ms += info["build_ms"]
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = ctx.kernel_launches() - launches0
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    agg = torch.tensor([float(launches), pair_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t.item())
    total_pairs = int(table.sum().item()) if world > 1 else int(local_rows.sum().item())
    pairs_local = int(local_rows.sum().item()) if world == 1 else int(table[rank * nf:(rank + 1) * nf].sum().item())
    value = total_pairs * args.steps / (elapsed_ms * 1e-3)

    # ---- e2e: the public API, evaluate.RDF over a trajectory of PINNED HOST arrays ------------
    # Every rank holds its own block of a global trajectory (ShardView); evaluate.RDF selects all
    # frames, takes this rank's contiguous block, streams it host -> device through the library's
    # staging pipeline, reads the histograms back, normalises, assembles the rows of all ranks
    # with one all-reduce and averages -- all inside the timed region.
    import amep_b200
    ef = nf if args.e2e_frames <= 0 else max(1, min(args.e2e_frames, nf))
    try:
        import psutil
        avail = psutil.virtual_memory().available
        # all ranks of this node pin memory at the same time: stay below a third of what is free
        ef = int(max(1, min(ef, avail / 3 / max(1, world) // (n * 12))))
    except Exception:
        pass
    host = torch.empty((ef, n, 3), dtype=torch.float32, pin_memory=True)
    host.copy_(coords[:ef])
    torch.cuda.synchronize()
    host_np = host.numpy()

    class ShardView(amep_b200.TensorTrajectory):
        """Global trajectory of ef * world frames of which this process holds [rank*ef, (rank+1)*ef)."""

        def __init__(self):
            super().__init__(host_np, box)
            self.steps = np.arange(ef * world)
            self.times = self.steps.astype(float)

        def __len__(self):
            return ef * world

        def stack(self, indices, ptype=None):
            idx = np.asarray(indices) - rank * ef
            assert len(idx) == 0 or (idx[0] >= 0 and idx[-1] < ef), "frame outside this rank's block"
            return super().stack(idx, ptype)

    traj = ShardView()

    def e2e_step():
        return amep_b200.evaluate.RDF(traj, skip=0.0, nav=ef * world, rmax=RMAX, nbins=NBINS)

    for _ in range(2):
        ev = e2e_step()
    # raw pinned host->device copy rate of the same buffer (the floor of any host-buffer path)
    scratch = torch.empty_like(coords[:ef])
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    scratch.copy_(host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = host.numel() * 4 / (h0.elapsed_time(h1) * 1e-3) / 1e9
    # the same with every rank copying at once: what the platform gives N concurrent streams
    barrier()
    h0.record()
    scratch.copy_(host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    tcc = torch.tensor([h0.elapsed_time(h1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tcc, op=dist.ReduceOp.MAX)
    h2d_gbs_concurrent = host.numel() * 4 / (float(tcc.item()) * 1e-3) / 1e9
    del scratch
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps)
    e2e_step_ms = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        ev = e2e_step()               # returns when frames / avg are complete on the host
        e2e_step_ms.append(round((time.perf_counter() - ts) * 1e3, 3))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_pairs = float(ev.counts.sum())                      # all ranks' frames (assembled table)
    e2e_value = e2e_pairs * e2e_steps / float(te.item())
    # both paths agree bit for bit
    np.testing.assert_array_equal(ev.counts[rank * ef:(rank + 1) * ef], local_rows[:ef].cpu().numpy())
    e2e_floor_s = ef * n * 12 / (h2d_gbs_concurrent * 1e9)
    del host, traj, ev

    del coords, table
    torch.cuda.empty_cache()
    # ---- strong scaling: BASELINE.json configs[4] ----------------------------------------
    strong = None
    if not args.no_strong:
        strong = strong_section(args, torch, dist, _core, ctx, dev, world, rank, barrier)

    # ---- S(q) ------------------------------------------------------------------------
    sq = None
    if not args.no_sq:
        sq = sq_section(args, torch, dist, _core, ctx, dev, world, rank, peaks_file())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the pair kernel -----------------------------------------------------
    peaks = peaks_file()
    ffma_peak = float(peaks.get("fp32_ffma_tflops", 69.2))                 # TFLOP/s, 2 flops per FFMA
    unfused_peak = float(peaks.get("fp32_mul_add_unfused_tflops", 35.9))   # TFLOP/s, 1 flop per instruction
    launches_pair = args.steps * info["sub_batches"]
    pair_ms_avg = pair_ms / max(1, launches_pair)
    achieved = pairs_local * args.steps * FLOPS_PER_PAIR_2D / (pair_ms * 1e-3) / 1e12 if pair_ms > 0 else 0.0
    roofline = {
        "bound": "fp32",
        "kernel": "amepb::pair_kernel<float,float,2,true,true> (float32 regime, z constant, shared tables, symmetric walk)",
        "achieved": achieved, "peak": ffma_peak, "unit": "TFLOP/s", "frac": achieved / ffma_peak,
        "peak_source": "FP32 FFMA throughput measured with tools/microbench.cu on this pool's B200 "
                       "(profiles/microbench_r01.json; theoretical 148 SM x 128 x 2 x 1.965 GHz = 74.5); "
                       "MEASURED_PEAKS.json has no CUDA-core peak (SURVEY.md 8d)",
        "frac_unfused": achieved / unfused_peak, "peak_unfused": unfused_peak,
        "peak_unfused_note": "FMUL+FADD instruction rate: the bit-exact arithmetic (every product rounded, no "
                             "contraction) cannot use FFMA for the distance itself",
        "flops_per_unit": FLOPS_PER_PAIR_2D, "units_per_launch": pairs_local / max(1, info["sub_batches"]),
        "avg_launch_ms": pair_ms_avg, "share_of_step": pair_ms / elapsed_ms,
        "candidate_pairs_per_s": info["candidate_pairs"] * args.steps / (pair_ms * 1e-3) if pair_ms > 0 else 0.0,
        "traffic": (peaks["pair_kernel_dram_bytes_per_frame"] * nf / max(1, info["sub_batches"])
                    if "pair_kernel_dram_bytes_per_frame" in peaks else None),
        "traffic_source": peaks.get("pair_kernel_dram_source"),
        "hbm_gbs_peak": hbm_peak(),
        "cell_build": {
            "bound": "hbm", "unit": "GB/s", "peak": hbm_peak(),
            "algorithmic_bytes_per_frame": 28 * n + 4 * info["cells"] / max(1, info["frames"]),
            "ms_per_frame": build_ms / args.steps / nf,
            "achieved": (28 * n + 4 * info["cells"] / max(1, info["frames"])) / (build_ms / args.steps / nf * 1e-3) / 1e9
            if build_ms > 0 else None,
            "note": "stats + count + scan + fill; 28 B/particle + 4 B/cell (SURVEY.md 8d)"},
    }
    if roofline["cell_build"]["achieved"]:
        roofline["cell_build"]["frac"] = roofline["cell_build"]["achieved"] / hbm_peak()

    cpu = None
    if not args.no_cpu and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import c_oracle as corc
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        nq = args.cpu_queries or auto_cpu_queries(n, threads)
        rate, dt, pairs, nimg = cpu_reference_rate(n, nq, threads)
        cpu = {"value": rate, "unit": "pair-evals/s", "cores": threads, "kind": "port",
               "sample": f"{nq} query particles of one frame x all {nimg} periodic images "
                         f"(brute force like amep/spatialcor.py:348-388), {dt:.1f} s"}

    line = {
        "metric": "rdf_pair_evals_per_s", "value": value, "unit": "pair-evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": "pair-evals/s", "h2d_bytes_per_step": ef * n * 12 * world,
                "d2h_bytes_per_step": ef * NBINS * 8 * world, "frames_per_gpu": ef,
                "api": "amep_b200.evaluate.RDF(TensorTrajectory over pinned host arrays, nav=all frames): H2D of "
                       "every frame, kernels, D2H of the histograms, NumPy normalisation, all-reduce of the rows, "
                       "frame average",
                "pinned_h2d_copy_gbs": h2d_gbs, "pinned_h2d_copy_gbs_all_ranks_at_once": h2d_gbs_concurrent,
                "h2d_floor_ms_per_frame": n * 12 / (h2d_gbs * 1e9) * 1e3,
                "platform_floor": {"value": e2e_pairs / max(e2e_floor_s, 1e-9), "unit": "pair-evals/s",
                                   "note": "pairs / time of the bare pinned H2D copies of the same frames with all "
                                           "ranks copying concurrently (no kernel, no epilogue)"},
                "frac_of_platform_floor": e2e_value / (e2e_pairs / max(e2e_floor_s, 1e-9)),
                "step_ms": e2e_step_ms},
        "gpu_launches": int(agg[0].item()),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "ms_per_frame": elapsed_ms / args.steps / nf, "pair_kernel_ms_per_frame": pair_ms / args.steps / nf,
        "build_ms_per_frame": build_ms / args.steps / nf,
        "cells_per_cutoff": info["cells_per_cutoff"], "strong": strong, "sq": sq,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def peaks_file():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "microbench_r01.json")))
    except Exception:
        return {}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0   # fallback of /opt/skills/guides/B200_PROFILING.md


def strong_section(args, torch, dist, _core, ctx, dev, world, rank, barrier):
    """BASELINE.json configs[4]: 3D uniform rho = 0.8, 2^24 particles, 64 frames in total, RDF rmax = 10,
    nbins = 500.  The frames are split over the ranks (contiguous blocks; fewer frames than ranks
    would shard inside the frame), one all-reduce of the [64, 500] histogram table per step."""
    from amep_b200 import dist as adist
    n, total = args.strong_particles, args.strong_frames
    rho = 0.8
    L = float((n / rho) ** (1.0 / 3.0))
    box = np.array([[0, L], [0, L], [0, L]], dtype=np.float32)
    edges = np.linspace(0.0, RMAX, NBINS + 1)
    intra = total < world
    lo, hi = (0, total) if intra else adist.shard_bounds(total, rank, world)
    nloc = hi - lo
    coords = torch.empty((max(nloc, 1), n, 3), device=dev, dtype=torch.float32)
    for f in range(nloc):      # frame f of the global trajectory has the same content on any rank count
        gen = torch.Generator(device=dev).manual_seed(50_000 + lo + f)
        coords[f] = (torch.rand((n, 3), generator=gen, device=dev, dtype=torch.float64) * L).float()
    table = torch.zeros((total, NBINS), dtype=torch.int64, device=dev)
    rows = table[lo:hi]

    def step():
        if world > 1:
            table.zero_()
        if nloc:
            _core.rdf_histograms(coords[:nloc], box, edges, out=rows, shard=(rank, world) if intra else None)
        if world > 1:
            dist.all_reduce(table, op=dist.ReduceOp.SUM)

    steps, warm = max(1, min(args.steps, 3)), 1
    for _ in range(warm):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pair_ms = build_ms = 0.0
    e0.record()
    for _ in range(steps):
        step()
        if nloc:
            info = ctx.rdf_info()
            pair_ms += info["pair_kernel_ms"]
            build_ms += info["build_ms"]
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    pairs = int(table.sum().item())
    checksum = int((table.sum(dim=0) * torch.arange(1, NBINS + 1, device=dev)).sum().item() % (1 << 61))
    expect = total * n * rho * 4.0 / 3.0 * np.pi * RMAX ** 3
    out = {"workload": f"BASELINE.json configs[4]: 3D uniform rho=0.8, N={n}, {total} frames in total split over "
                       f"{world} GPU(s), RDF rmax=10, nbins=500, float32 regime",
           "scaling": "strong", "value": pairs / (ms * 1e-3), "unit": "pair-evals/s", "ms_per_step": ms,
           "steps": steps, "warmup": warm, "frames_total": total, "frames_this_rank": nloc,
           "pairs_per_step": pairs, "ideal_gas_expectation": expect, "rel_dev_from_ideal_gas": (pairs - expect) / expect,
           "histogram_checksum": checksum,
           "pair_kernel_ms_per_frame_rank0": pair_ms / steps / max(1, nloc), "build_ms_per_frame_rank0": build_ms / steps / max(1, nloc),
           "flops_per_pair": 8,
           "pair_kernel_tflops_rank0": (pairs / total * nloc * 8 / (pair_ms / steps * 1e-3) / 1e12) if pair_ms > 0 else None}
    del coords, table
    torch.cuda.empty_cache()
    return out


def sq_cpu_baselines(threads):
    """The reference's S(q) inner loops (C restatement in oracle/rdf_oracle.c) on the host cores, on
    bounded samples: __sf2d_chunk_std (amep/spatialcor.py:1690-1719, complex64 running sums) on a
    slice of the configs[3] grid, and __sf_iso_chunk_fast (amep/spatialcor.py:1402-1448) on one
    direction of the 3D case."""
    import ctypes
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle as corc
    out = {}
    rng = np.random.default_rng(4)
    # sf2d: 65 536 particles of configs[3] against a slice of the 512 x 512 grid
    n, L = 65_536, 512.0
    c = np.zeros((n, 3), dtype=np.float32)
    c[:, :2] = rng.uniform(0, L, (n, 2)).astype(np.float32)
    box = np.array([[0, L], [0, L], [-0.5, 0.5]], dtype=np.float32)
    qmax = 256.5 * 2 * np.pi / 512.0
    npts = 512 * threads
    t0 = time.perf_counter()
    corc.sf2d_modes(c, box, qmax, 715, nthreads=threads, grid_slice=(0, npts))
    dt = time.perf_counter() - t0
    if dt < 3.0:     # aim at a few seconds
        npts = int(min(512 * 512, npts * 4.0 / max(dt, 1e-3)))
        npts = max(threads, npts // threads * threads)
        t0 = time.perf_counter()
        corc.sf2d_modes(c, box, qmax, 715, nthreads=threads, grid_slice=(0, npts))
        dt = time.perf_counter() - t0
    out["sf2d_std"] = {"value": n * npts / dt, "unit": "particle*q terms/s", "cores": threads, "kind": "port",
                       "sample": f"{n} particles x {npts} points of the 512x512 grid of configs[3], complex64 running sums "
                                 f"in the reference's order (sincos in float64 per term), {dt:.1f} s"}
    # sfiso fast: one direction, 544 q values, 3D
    n3 = 262_144
    L3 = float((n3 / 0.8) ** (1 / 3))
    c3 = rng.uniform(0, L3, (n3, 3)).astype(np.float32)
    q = np.ascontiguousarray(np.arange(1, 545) * (2 * np.pi / 171.0))
    u = np.ascontiguousarray(np.array([0.6, 0.0, 0.8]))
    sums = np.empty((len(q), 2))
    dp = ctypes.POINTER(ctypes.c_double)
    ndir = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 3.0 and ndir < 32:
        corc.lib().oracle_sfiso_sums(c3.ctypes.data, 0, n3, u.ctypes.data_as(dp), q.ctypes.data_as(dp), len(q),
                                     sums.ctypes.data_as(dp), threads)
        ndir += 1
    dt = time.perf_counter() - t0
    out["sfiso_fast"] = {"value": n3 * len(q) * ndir / dt, "unit": "particle*q terms/s", "cores": threads, "kind": "port",
                         "sample": f"{n3} particles x {len(q)} q values x {ndir} direction(s) (float64 cos and sin per term, as "
                                   f"amep/spatialcor.py:1440-1444), {dt:.1f} s"}
    return out


def sq_section(args, torch, dist, _core, ctx, dev, world, rank, peaks):
    """S(q) throughput: particle*q terms per second, reference-equivalent term counts
    (SURVEY.md 8d): sfiso fast = N * nq * n_directions; sf2d = N * nx * ny."""
    from amep_b200 import spatialcor as sc
    out = {}
    gen = torch.Generator(device=dev).manual_seed(99 + rank)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        t = torch.tensor([best], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # sfiso(mode='fast'), 3D LJ-like density (configs[2] shape at reduced N): rho=0.8, qmax=20, num=8
    n3, nf3 = 1_000_000, 4
    L3 = float((n3 / 0.8) ** (1 / 3))
    c3 = torch.rand((nf3, n3, 3), generator=gen, device=dev, dtype=torch.float32) * L3
    box3 = np.array([[0, L3], [0, L3], [0, L3]], dtype=np.float32)
    res = {}

    def run_sfiso():
        res["s"], res["q"] = sc.sfiso_frames(c3, box3, None, qmax=20.0, twod=False, mode="fast", num=8)
    ms = timed(run_sfiso)
    nq = res["q"].shape[1]
    dfma_instr_peak = float(peaks.get("fp64_dfma_tflops", 33.9)) / 2 * 1e12     # FP64-pipe instructions per second

    def fp64_roofline(kernel, instr, kernel_ms, what):
        rate = instr / (kernel_ms * 1e-3) if kernel_ms > 0 else 0.0
        return {"bound": "fp64 pipe", "kernel": kernel, "achieved": rate / 1e12, "peak": dfma_instr_peak / 1e12,
                "unit": "T FP64-pipe instructions/s", "frac": rate / dfma_instr_peak, "instructions_per_launch_set": instr,
                "counted": what, "kernel_ms": kernel_ms,
                "peak_source": "DFMA rate measured with tools/microbench.cu (profiles/microbench_r01.json: "
                               "33.9 TFLOP/s = 16.9 T DFMA/s; theoretical 148 x 64 x 1.965 GHz = 18.6 T/s)"}

    kms = ctx.last_kernel_ms()
    out["sfiso_fast"] = {"workload": f"3D uniform rho=0.8, N={n3}, qmax=20, num=8 (nq={nq}, 32 directions, 13 unique), {nf3} frames per GPU",
                         "terms_per_s": nf3 * n3 * nq * 32 * world / (ms * 1e-3), "ms": ms, "kernel_ms": kms,
                         "unique_terms_per_s": nf3 * n3 * nq * 13 * world / (ms * 1e-3), "dtype": "f64",
                         "roofline": fp64_roofline("amepb::harmonics_kernel<float>", nf3 * n3 * nq * 13 * 4.0, kms,
                                                   "2 DFMA + 2 DADD per (particle, harmonic, unique direction): Chebyshev "
                                                   "recurrence for cos and sin")}
    del c3
    # sf2d(mode='std'), configs[3]: 256k particles, 512 x 512 q grid
    n2, L2, nf2 = 262_144, 512.0, 8
    c2 = torch.zeros((nf2, n2, 3), device=dev, dtype=torch.float32)
    c2[:, :, :2] = torch.rand((nf2, n2, 2), generator=gen, device=dev, dtype=torch.float32) * L2
    box2 = np.array([[0, L2], [0, L2], [-0.5, 0.5]], dtype=np.float32)
    qmax2 = 256.5 * 2 * np.pi / 512.0
    for accum in ("c64", "f64"):
        def run_sf2d():
            res["s2"] = sc.sf2d_frames(c2, box2, qmax=qmax2, accum=accum)[0]
        ms = timed(run_sf2d, reps=2)
        g = res["s2"].shape[1]
        kms = ctx.last_kernel_ms()
        per_point = 2.0 if accum == "c64" else 1.0
        out["sf2d_std_" + accum] = {
            "workload": f"configs[3]: 2D uniform N={n2}, {g}x{g} q grid, {nf2} frames per GPU, "
                        + ("complex64 accumulation in the reference's order" if accum == "c64" else "float64 accumulation"),
            "terms_per_s": nf2 * n2 * g * g * world / (ms * 1e-3), "ms": ms, "kernel_ms": kms,
            "dtype": "f64 (rounded to f32 per add)" if accum == "c64" else "f64",
            "roofline": fp64_roofline("amepb::sf2d_c64_kernel<float>" if accum == "c64" else "amepb::sf2d_f64_kernel<float>",
                                      nf2 * n2 * g * g * per_point, kms,
                                      "8 DFMA per particle and quadruple of grid points (+-qx, +-qy) = 2 per grid point: every "
                                      "running sum is rounded to float32 after each term, so real and imaginary parts of both "
                                      "sign combinations are separate chains" if accum == "c64" else
                                      "4 DFMA per particle and quadruple of grid points = 1 per grid point (separable "
                                      "cos/sin products; SURVEY's 8 flops per unique term overstate this formulation)")}
    del c2
    # pcf2d (widened row): 2D, density 1, 500 x 500 grid; the reference's default
    # window (half of all pairs bin) with an orientation, and a narrow window.  Never allowed to
    # break the headline line.
    try:
        npc = 100_000
        Lp = float(np.sqrt(npc))
        cp = torch.zeros((1, npc, 3), device=dev, dtype=torch.float32)
        # box centred on the origin: the reference rotates the difference vectors about the box centre
        cp[:, :, :2] = (torch.rand((1, npc, 2), generator=gen, device=dev, dtype=torch.float32) - 0.5) * Lp
        boxp = np.array([[-Lp / 2, Lp / 2], [-Lp / 2, Lp / 2], [-0.5, 0.5]], dtype=np.float32)
        for tag, kw in (("default_window_rotated", dict(psi=np.array([0.8, 0.3]))), ("rmax10", dict(rmax=10.0))):
            def run_pcf2d():
                res["g"] = sc.pcf2d_frames(cp, boxp, zero_last_column=False, **kw)[0]
            # timed on this rank only (no collective inside the guarded block: a rank that fails
            # here must not leave the others waiting); every rank does the same work
            run_pcf2d()
            torch.cuda.synchronize()
            ms = 1e30
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run_pcf2d()
                e1.record()
                torch.cuda.synchronize()
                ms = min(ms, e0.elapsed_time(e1))
            area = (2.0 * (Lp // (2 * np.sqrt(2)) if "rmax" not in kw else kw["rmax"]) / 500) ** 2
            binned = float(np.sum(res["g"][0]) * area * npc)   # undo the normalisation: g = hist / area / (N / L^2) / N
            out["pcf2d_" + tag] = {
                "workload": f"2D uniform N={npc}, density 1, 500x500 (dx, dy) grid, 1 frame per GPU, " + tag,
                "pair_evals_per_s": npc * (npc - 1) * world / (ms * 1e-3), "binned_pairs_per_s": binned * world / (ms * 1e-3),
                "ms": ms, "dtype": "f32 differences, exact np.histogram2d binning"}
    except Exception as exc:  # noqa: BLE001
        out["pcf2d"] = {"error": repr(exc)}
    # pcf_angle / spatialcor (remaining functions): all-pairs kernels, one frame
    try:
        npa = 100_000
        Lp = float(np.sqrt(npa))
        cpa = torch.zeros((npa, 3), device=dev, dtype=torch.float32)
        cpa[:, :2] = torch.rand((npa, 2), generator=gen, device=dev, dtype=torch.float32) * Lp
        boxa = np.array([[0, Lp], [0, Lp], [-0.5, 0.5]], dtype=np.float32)
        vel = torch.randn((npa, 3), generator=gen, device=dev, dtype=torch.float32)

        def one(fn):
            fn()
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best
        for tag, kw in (("rmax10", dict(rmax=10.0)), ("default_rmax", dict())):
            ms = one(lambda: sc.pcf_angle(cpa, boxa, psi=np.array([0.8, 0.3]), **kw))
            out["pcf_angle_" + tag] = {"workload": f"2D uniform N={npa}, density 1, 500 x 100 (r, theta) grid, psi rotation, {tag}",
                                       "pair_evals_per_s": npa * (npa - 1) * world / (ms * 1e-3), "ms": ms,
                                       "dtype": "f32 distances (exact fold / norm), f64 angles"}
        ms = one(lambda: sc.spatialcor(cpa, boxa, vel))
        out["spatialcor_velocity"] = {"workload": f"2D uniform N={npa}, density 1, 3-component float32 values, default bins "
                                                  "(1.05 wide up to L/2)", "pair_evals_per_s": npa * npa * world / (ms * 1e-3),
                                      "ms": ms, "dtype": "f64 minimum-image distances and sums"}
    except Exception as exc:  # noqa: BLE001
        out["pair_observables"] = {"error": repr(exc)}
    if world == 1 and rank == 0 and not args.no_cpu:
        try:
            threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            out["cpu_baseline"] = sq_cpu_baselines(threads)
        except Exception as exc:  # noqa: BLE001
            out["cpu_baseline"] = {"error": repr(exc)}
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
