#!/usr/bin/env python3
"""bench.py — throughput of the streamline-metrics hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--streamlines S]

A "step" is one pass of the hot path over one synthetic tractogram already resident in HBM:
tg_metrics_csr_dev (17 metrics per polyline) + tg_bundle_reduce_dev (bundle partial moments)
[+ an NCCL all-gather of the 27 bundle partials per rank when N > 1], all through the C ABI of
include/tractgeom.h.  Workload at N=1 = BASELINE.json configs[4] restricted to one GPU, the
configuration the 60 %-of-HBM target is quoted on: 10M polylines, n = clip(round(N(100,15^2)),3,200),
~1e9 points, 24 GB of float64 coordinates (>> the 126 MB L2, so no flush between iterations).
At N > 1 every rank owns its own 10M-polyline CSR shard of an N x 10M tractogram (weak scaling).

One JSON line on stdout (rank 0).  `value` = polylines/s over all ranks, device-resident;
`e2e` = the same metric through the HOST-buffer call tg_metrics_csr_host (pinned host buffers,
H2D + kernels + D2H inside the timed region); `roofline` = algorithmic bytes of the metrics kernel
/ its CUDA-event time vs MEASURED_PEAKS.json; `cpu_baseline` = the numpy oracle port on the host cores.

--impl reference times the CPU implementation (oracle port of the reference's numpy path; the
Python reference itself cannot travel to the GPU box) on all host cores, same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "streamlines_per_sec"
UNIT = "streamlines/s"
CFG_MEAN, CFG_SD, CFG_LO, CFG_HI = 100.0, 15.0, 3, 200


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streamlines", type=int, default=int(os.environ.get("TG_BENCH_STREAMLINES", 10_000_000)),
                    help="polylines per GPU (default: 10M = BASELINE config 5 on one GPU)")
    ap.add_argument("--law", default="normal", choices=["normal", "heavy", "loguniform", "fixed96"],
                    help="length law (heavy = config 4; loguniform = its long-polyline variant; fixed96 = traffic probe)")
    ap.add_argument("--e2e-streamlines", type=int, default=int(os.environ.get("TG_BENCH_E2E_STREAMLINES", 1_000_000)))
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("TG_BENCH_CPU_SAMPLE", 100_000)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="timing experiments with a deliberately incomplete kernel")
    return ap.parse_args()


def workload_name(S, law):
    if law == "normal":
        return f"synthetic tractogram, {S} polylines/GPU x ~100 points (n=clip(round(N(100,15^2)),3,200)), float64 CSR (BASELINE configs[4])"
    if law == "loguniform":
        return f"log-uniform tractogram, {S} polylines/GPU, n=floor(10*500^U) in [10,5000], mean ~803 (variant of BASELINE configs[3])"
    return f"heavy-tailed tractogram, {S} polylines/GPU, n=min(5000,floor(10/U)) (BASELINE configs[3])"


def bind_to_gpu_numa_node(local):
    """One process per GPU: run on the cores of the NUMA node the GPU hangs off, so that the pinned host buffers of
    the e2e leg are allocated next to it (cudaHostAlloc places pages by first touch).  Best effort; returns the node."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def algorithmic_bytes(P, S):
    """BASELINE.md §5 / SURVEY.md §8d."""
    return 24 * P + 8 * (S + 1) + 136 * S + S


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.f = None
        self.p = None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power) if power else None)
        return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle port; test infrastructure used only as the reported baseline)
# ------------------------------------------------------------------------------------------------
def host_sample(S, law, seed=5):
    """The first S polylines' worth of the workload's length law, generated on the host (same curve model)."""
    import numpy as np
    from lesion_condition_vae_b200 import synth
    rng = np.random.default_rng(seed)
    if law == "normal":
        n = synth.lengths_normal(rng, S, CFG_MEAN, CFG_SD, CFG_LO, CFG_HI)
    elif law == "loguniform":
        n = synth.lengths_log_uniform(rng, S)
    else:
        n = synth.lengths_heavy_tail(rng, S)
    return synth.random_walk_csr(n, seed)


def cpu_leg(pts, off, cores):
    from oracle import cpu_bench
    return cpu_bench.timed_run(pts, off, cores)


def run_reference(args):
    """--impl reference: the CPU path on all host cores, same metric/config; each step = one bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = max(cores * 250, 2000)            # ~0.3 ms per polyline per core -> ~0.1-0.3 s per step
    pts, off = host_sample(per_step, args.law)
    from oracle import cpu_bench
    pool = cpu_bench.Pool(cores)
    try:
        for _ in range(args.warmup):
            pool.run(pts, off)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            rows = pool.run(pts, off)
        dt = time.perf_counter() - t0
    finally:
        pool.close()
    assert rows == per_step, (rows, per_step)
    value = per_step * args.steps / dt
    P = int(off[-1])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.streamlines, args.law), "sample": f"{per_step} polylines ({P} points) per step"},
        "points_per_sec": P * args.steps / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} polylines x {args.steps} steps, numpy oracle port of tract_geom_proc.py:153-212, one process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    from lesion_condition_vae_b200 import _lib, sharding, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU implementation (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None      # before any pinned allocation (first touch)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    ctx = _lib.Context(local)
    S = args.streamlines
    seed = 5 + 1000 * rank
    n = synth.torch_lengths(args.law, S, seed, dev)
    pts, off = synth.torch_random_walk_csr(n, seed, dev)
    P = int(pts.shape[0])
    del n
    out = torch.empty((17, S), dtype=torch.float64, device=dev)
    keep = torch.empty(S, dtype=torch.uint8, device=dev)
    part = torch.zeros((1, sharding.PARTIAL_WIDTH), dtype=torch.float64, device=dev)   # 13 sums | 14 counts (as f64): all-gather payload
    sums = torch.empty((1, 13), dtype=torch.float64, device=dev)
    counts = torch.empty((1, 14), dtype=torch.int64, device=dev)
    bo = np.array([0, S], dtype=np.int64)
    stream = torch.cuda.Stream(dev)           # a real (non-default) stream: kernels AND timing events live on it
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def step(ev=None):
        if ev is not None:
            ev[0].record(stream)
        ctx.metrics_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, out.data_ptr(), keep.data_ptr(), sp)
        if ev is not None:
            ev[1].record(stream)
        ctx.bundle_reduce_dev(out.data_ptr(), keep.data_ptr(), 0, S, bo, sums.data_ptr(), counts.data_ptr(), sp)
        if world > 1:
            part[0, :13] = sums[0]
            part[0, 13:] = counts[0].to(torch.float64)           # exact: counts < 2^53
            return sharding.allgather_partials(part)             # (world, 1, 27): summed in rank order by the caller
        return None

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize(dev)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    launches0 = ctx.launches
    e0.record(stream)
    for i in range(args.steps):
        step(kev[i])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    launches = ctx.launches - launches0
    clk = clocks.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    k_ms = sorted(a.elapsed_time(b) for a, b in kev)
    k_avg = sum(k_ms) / len(k_ms)
    tot_S, tot_P = float(S), float(P)
    if world > 1:
        t = torch.tensor([ms_total, k_avg], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, k_avg_max = float(t[0]), float(t[1])
        c = torch.tensor([tot_S, tot_P, float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_S, tot_P, launches = float(c[0]), float(c[1]), int(c[2])
    else:
        k_avg_max = k_avg
    sec = ms_total / 1e3
    value = tot_S * args.steps / sec

    # sanity: the result is real (every polyline kept, finite lengths); at N > 1 from the gathered partials
    if world > 1:
        g_sums, g_counts = sharding.combine_partials(step().cpu().numpy())
        n_kept = int(g_counts[0, 0]); mean_len = float(g_sums[0, 0]) / max(int(g_counts[0, 1]), 1)
    else:
        n_kept = int(counts[0, 0].item())
        mean_len = float(sums[0, 0].item()) / max(n_kept, 1)

    # ---- e2e: HOST buffers through tg_metrics_csr_host (pinned), H2D + kernels + D2H per step ----
    Se = min(args.e2e_streamlines, S)
    Pe = int(off[Se].item())
    h_pts = torch.empty((Pe, 3), dtype=torch.float64, pin_memory=True)
    h_off = torch.empty(Se + 1, dtype=torch.int64, pin_memory=True)
    h_pts.copy_(pts[:Pe]); h_off.copy_(off[:Se + 1])
    torch.cuda.synchronize(dev)
    hp, ho = h_pts.numpy(), h_off.numpy()
    h_out = torch.empty((17, Se), dtype=torch.float64, pin_memory=True).numpy()      # result buffers the caller owns, pinned
    h_keep = torch.empty(Se, dtype=torch.uint8, pin_memory=True).numpy()
    e2e_steps = max(3, min(args.steps, 10))
    ctx.metrics_host(hp, ho, out=h_out, keep=h_keep)             # warm-up (allocates device scratch)
    ctx.metrics_host(hp, ho, out=h_out, keep=h_keep)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        o_h, k_h, s_h, c_h = ctx.metrics_host(hp, ho, out=h_out, keep=h_keep)
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t[0])
    e2e_value = world * Se * e2e_steps / e2e_sec
    h2d = 24 * Pe + 8 * (Se + 1)
    d2h = 17 * 8 * Se + Se + 13 * 8 + 14 * 8
    assert args.no_check or (int(c_h[0, 0]) == Se and np.isfinite(o_h[0]).all())
    # same call with the points stored as float32 on the host (what legacy-VTK tract files hold; upcast exactly on
    # the device): half the H2D bytes.  Reported beside `e2e`, never instead of it.
    h_pts32 = torch.empty((Pe, 3), dtype=torch.float32, pin_memory=True)
    h_pts32.copy_(pts[:Pe])
    torch.cuda.synchronize(dev)
    hp32 = h_pts32.numpy()
    ctx.metrics_host(hp32, ho, out=h_out, keep=h_keep)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.metrics_host(hp32, ho, out=h_out, keep=h_keep)
    e2e32_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e32_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e32_sec = float(t[0])
    del h_pts32, hp32

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peaks()
    abytes = algorithmic_bytes(P, S)
    achieved = abytes / (k_avg_max * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("streamlines") == S and tj.get("law") == args.law:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(S, args.law), "streamlines_per_gpu": S, "points_per_gpu": P,
                   "parallelism": f"csr-range shards x{world}" + (", NCCL all-gather of 27 bundle partials/rank" if world > 1 else "")
                   + (f", ranks bound to their GPU's NUMA node (rank 0: node {numa})" if numa is not None else ""),
                   "l2": "inputs (24 B/point, >> 126 MB L2) stream from HBM every step; no flush needed",
                   "e2e_workload": f"first {Se} polylines ({Pe} points) per GPU from pinned host buffers, full df_sl table copied back"},
        "points_per_sec": tot_P * args.steps / sec,
        "kernel_ms": {"metrics_avg": k_avg_max, "metrics_min": k_ms[0], "metrics_median": k_ms[len(k_ms) // 2]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes,
                     "kernel": "k_metrics (17 metrics per polyline), avg CUDA-event time over the timed steps",
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        # the unit that actually binds the kernel: 96 fp64-pipe instructions per point (SASS of the
        # steady block, profiles/r1_grouped_blocks.txt) against 60 lane-operations / clock / SM (tools/microbench.cu)
        "fp64_pipe": {"lane_ops_per_point": 96, "achieved_tlaneops": 96 * P / (k_avg_max * 1e-3) / 1e12,
                      "peak_tlaneops_at_max_clock": 60 * 148 * 1.965e9 / 1e12,
                      "frac": 96 * P / (k_avg_max * 1e-3) / (60 * 148 * 1.965e9)},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_sec / e2e_steps, "steps": e2e_steps},
        "e2e_f32_points": {"value": world * Se * e2e_steps / e2e32_sec, "unit": UNIT, "h2d_bytes_per_step": 12 * Pe + 8 * (Se + 1),
                           "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e32_sec / e2e_steps,
                           "note": "same call, host points stored as float32 (legacy-VTK float files), computed in fp64"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "check": {"n_streamlines": n_kept, "length_mean": mean_len},
    }
    if not args.no_cpu_baseline and world == 1:
        cs = min(args.cpu_sample, Se)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        cp = hp[:int(ho[cs])].copy(); co = ho[:cs + 1].copy()
        rows, dt = cpu_leg(cp, co, cores)
        line["cpu_baseline"] = {"value": rows / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {cs} polylines ({int(co[-1])} points) of the same tractogram, numpy oracle port, one process per core, {dt:.1f} s"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
