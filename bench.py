#!/usr/bin/env python
"""Benchmark of the RDF / S(q) hot path (BASELINE.json metric) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores

A *step* is one pass of the hot path over one batch of synthetic frames.
Workload at every N (weak scaling, frames sharded over ranks, one NCCL
all-reduce of the per-frame histograms): BASELINE.json configs[1], "2D active
Brownian particles, 1M particles, RDF rmax=10 sigma, nbins=500" -- uniform
positions at packing fraction 0.5 (rho = 0.6366, L = 1253.3), `--frames` frames
per GPU (default 1000, the config's count; every frame is the same amount of work
repeated).

One JSON line on stdout (rank 0):
  value      in-range ordered pair evaluations per second (sum of all histogram
             counts / time), inputs resident in HBM, all ranks, all-reduce included
  e2e        the same metric through the C ABI with HOST (pinned) buffers:
             host->device copies of the coordinates and device->host of the
             histograms inside the timed region
  roofline   the pair kernel (dominant launch): algorithmic flops (5 per in-range
             2D pair: 2 SUB, 2 MUL, 1 ADD -- SURVEY.md 8d) over its CUDA-event time,
             against the FP32 non-FMA instruction peak measured by tools/microbench
  cpu_baseline  the C restatement of the reference algorithm (oracle/rdf_oracle.c,
             brute force over all periodic images like amep/spatialcor.py:348-388)
             on the host cores, on a bounded sample of the same frame
  sq         S(q) particle*q terms per second for sfiso(mode='fast') and sf2d(mode='std')
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
    ap.add_argument("--e2e-frames", type=int, default=256, help="frames per GPU of the host-buffer leg")
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

    # ---- e2e: host (pinned) buffers through the C ABI ----------------------------------
    ef = max(1, min(args.e2e_frames, nf))
    host = torch.empty((ef, n, 3), dtype=torch.float32).pin_memory()
    host.copy_(coords[:ef])
    host_np = host.numpy()
    hist_host = torch.empty((ef, NBINS), dtype=torch.int64).pin_memory()
    hist_np = hist_host.numpy()
    for _ in range(2):
        _core.rdf_histograms(host_np, box, edges, out=hist_np, device=local)
    # raw pinned host->device copy rate of the same buffer (the floor of any host-buffer path)
    scratch = torch.empty_like(coords[:ef])
    torch.cuda.synchronize()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    scratch.copy_(host, non_blocking=True)
    h1.record()
    torch.cuda.synchronize()
    h2d_gbs = host.numel() * 4 / (h0.elapsed_time(h1) * 1e-3) / 1e9
    del scratch
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps)
    e2e_step_ms = []
    for _ in range(e2e_steps):
        ts = time.perf_counter()
        _core.rdf_histograms(host_np, box, edges, out=hist_np, device=local)   # returns when hist_np is complete
        e2e_step_ms.append(round((time.perf_counter() - ts) * 1e3, 3))
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    pe = torch.tensor([float(hist_np.sum())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(pe, op=dist.ReduceOp.SUM)
    e2e_value = float(pe.item()) * e2e_steps / float(te.item())
    np.testing.assert_array_equal(hist_np, local_rows[:ef].cpu().numpy())   # both paths agree bit for bit
    del host, hist_host

    # ---- S(q) ------------------------------------------------------------------------
    sq = None
    if not args.no_sq:
        sq = sq_section(torch, dist, _core, ctx, dev, world, rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the pair kernel -----------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "microbench_r01.json")))
    except Exception:
        pass
    fp32_peak = float(peaks.get("fp32_mul_add_unfused_tflops", 35.9))   # TFLOP/s, 1 flop per instruction
    launches_pair = args.steps * info["sub_batches"]
    pair_ms_avg = pair_ms / max(1, launches_pair)
    achieved = pairs_local * args.steps * FLOPS_PER_PAIR_2D / (pair_ms * 1e-3) / 1e12 if pair_ms > 0 else 0.0
    roofline = {
        "bound": "fp32",
        "kernel": "amepb::pair_kernel<float,float,2,true,true> (float32 regime, z constant, shared tables, symmetric walk)",
        "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
        "peak_source": "tools/microbench.cu on this pool's B200 (profiles/microbench_r01.json): FP32 FMUL+FADD "
                       "(non-fused, the bit-exact path may not contract) instruction throughput; "
                       "MEASURED_PEAKS.json has no CUDA-core peak",
        "flops_per_unit": FLOPS_PER_PAIR_2D, "units_per_launch": pairs_local / max(1, info["sub_batches"]),
        "avg_launch_ms": pair_ms_avg, "share_of_step": pair_ms / elapsed_ms,
        "candidate_pairs_per_s": info["candidate_pairs"] * args.steps / (pair_ms * 1e-3) if pair_ms > 0 else 0.0,
        "traffic": (peaks["pair_kernel_dram_bytes_per_frame"] * nf / max(1, info["sub_batches"])
                    if "pair_kernel_dram_bytes_per_frame" in peaks else None),
        "traffic_source": peaks.get("pair_kernel_dram_source"),
        "hbm_gbs_peak": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
        if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0,
    }

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
                "pinned_h2d_copy_gbs": h2d_gbs,
                "h2d_floor_ms_per_frame": n * 12 / (h2d_gbs * 1e9) * 1e3, "step_ms": e2e_step_ms},
        "gpu_launches": int(agg[0].item()),
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "ms_per_frame": elapsed_ms / args.steps / nf, "pair_kernel_ms_per_frame": pair_ms / args.steps / nf,
        "build_ms_per_frame": build_ms / args.steps / nf,
        "cells_per_cutoff": info["cells_per_cutoff"], "sq": sq,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sq_section(torch, dist, _core, ctx, dev, world, rank):
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
    out["sfiso_fast"] = {"workload": f"3D uniform rho=0.8, N={n3}, qmax=20, num=8 (nq={nq}, 32 directions, 13 unique), {nf3} frames per GPU",
                         "terms_per_s": nf3 * n3 * nq * 32 * world / (ms * 1e-3), "ms": ms, "kernel_ms": ctx.last_kernel_ms(),
                         "unique_terms_per_s": nf3 * n3 * nq * 13 * world / (ms * 1e-3), "dtype": "f64"}
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
        out["sf2d_std_" + accum] = {
            "workload": f"configs[3]: 2D uniform N={n2}, {g}x{g} q grid, {nf2} frames per GPU, "
                        + ("complex64 accumulation in the reference's order" if accum == "c64" else "float64 accumulation"),
            "terms_per_s": nf2 * n2 * g * g * world / (ms * 1e-3), "ms": ms, "kernel_ms": ctx.last_kernel_ms(),
            "dtype": "f64 (rounded to f32 per add)" if accum == "c64" else "f64"}
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
    return out


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
