#!/usr/bin/env python3
"""bench.py — throughput of the streamline-metrics hot path on B200 (driver contract).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 0..4] [--scaling weak|strong]

A "step" is one pass of the hot path over one synthetic tractogram already resident in HBM:
tg_metrics_csr_dev (17 metrics per polyline) + tg_bundle_partials_dev (bundle partial moments, kernel 2 writes
the 27-double exchange payload) [+ ONE NCCL all-gather of that payload when N > 1], all through the C ABI of
include/tractgeom.h.  Default workload = BASELINE.json configs[4] restricted to one GPU, the configuration the
60 %-of-HBM target is quoted on: 10M polylines, n = clip(round(N(100,15^2)),3,200), ~1e9 points, 24 GB of float64
coordinates (>> the 126 MB L2, so no flush between iterations).  `--config k` selects BASELINE configs[k].

Scaling.  The headline line is WEAK (every rank owns its own 10M-polyline tractogram; `value` stays comparable
across N).  At N > 1 the line also carries `strong`: ONE 10M-polyline tractogram (the same on every rank, seeded)
cut into CSR ranges by sharding.shard_ranges (balanced by points), each rank running its slice — north_star's
"shard by CSR range" of BASELINE configs[4].  `--scaling strong` makes that the headline instead.

One JSON line on stdout (rank 0).  `value` = polylines/s over all ranks, device-resident; `e2e` = the same metric
through the HOST-buffer call tg_metrics_csr_host (pinned host buffers, H2D + kernels + D2H inside the timed
region); `e2e_file` (N = 1) = file -> two DataFrames through compute_streamline_metrics / the batched driver for
configs[0] / configs[1]; `roofline` = algorithmic bytes of the metrics call / its CUDA-event time vs
MEASURED_PEAKS.json; `cpu_baseline` = the numpy oracle port on the host cores.

--impl reference times the CPU implementation (oracle port of the reference's numpy path; the Python reference
itself cannot travel to the GPU box) on all host cores, same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "streamlines_per_sec"
UNIT = "streamlines/s"

# BASELINE.json configs[k] (SURVEY.md §8d "Config 1..5"): length law, polylines, seed, bundles
CONFIGS = {
    0: dict(law="uniform", S=1_000, seed=0, bundles=1,
            name="configs[0]: one tract of 1,000 polylines, n ~ U{20..119} (max_streamlines=1000)"),
    1: dict(law="uniform40", S=64 * 5_000, seed=1000, bundles=64,
            name="configs[1]: 16 tracts x 4 timepoints = 64 bundles x 5,000 polylines, n ~ U{40..139}, one batched call"),
    2: dict(law="normal", S=1_000_000, seed=3, bundles=1,
            name="configs[2]: whole-brain tractogram, 1M polylines x ~100 points"),
    3: dict(law="heavy", S=2_000_000, seed=4, bundles=1,
            name="configs[3]: heavy-tailed ragged lengths, 2M polylines, n = min(5000, floor(10/U))"),
    4: dict(law="normal", S=10_000_000, seed=5, bundles=1,
            name="configs[4]: 10M polylines x ~100 points (n = clip(round(N(100,15^2)),3,200)), ~1e9 points"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=None, choices=sorted(CONFIGS), help="BASELINE.json configs[k] (default: 4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--streamlines", type=int, default=int(os.environ.get("TG_BENCH_STREAMLINES", 0)) or None,
                    help="polylines per GPU (overrides the config's count)")
    ap.add_argument("--law", default=None, choices=["normal", "heavy", "loguniform", "fixed96", "uniform", "uniform40"],
                    help="length law (overrides the config's)")
    ap.add_argument("--e2e-streamlines", type=int, default=int(os.environ.get("TG_BENCH_E2E_STREAMLINES", 0)) or None,
                    help="polylines of the e2e leg (default: the whole workload at N=1 when host memory allows)")
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("TG_BENCH_CPU_SAMPLE", 100_000)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e-file", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record at N > 1")
    ap.add_argument("--no-check", action="store_true", help="timing experiments with a deliberately incomplete kernel")
    a = ap.parse_args()
    cfg = dict(CONFIGS[4 if a.config is None else a.config])
    if a.law:
        cfg["law"] = a.law
        cfg["name"] = f"{a.law} length law (variant of {cfg['name'].split(':')[0]})"
    if a.streamlines:
        cfg["S"] = a.streamlines
    a.cfg = cfg
    return a


def workload_name(cfg):
    return f"synthetic tractogram, BASELINE {cfg['name']}; {cfg['S']} polylines/GPU, float64 CSR"


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


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 0


# ------------------------------------------------------------------------------------------------
# clocks: NVML polled in-process (a thread, every 20 ms, started >= 1 s before the timed region so that even a
# 0.2 s region holds samples); nvidia-smi as the fallback
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x2: "applications_clocks_setting", 0x10: "sync_boost", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        self.index = index
        self.samples = []            # (t, sm_mhz, reasons bitmask, power W)
        self.stop_flag = threading.Event()
        self.thread = None
        self.max_mhz = None
        self.t0 = self.t1 = None
        self.fallback = None

    def _run(self, nv, h):
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.samples.append((time.perf_counter(), float(sm), int(rs), pw))
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # NVML enumerates in PCI order; CUDA_VISIBLE_DEVICES may remap: match by PCI bus id
            import torch
            pr = torch.cuda.get_device_properties(self.index)
            bdf = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            try:
                h = nv.nvmlDeviceGetHandleByPciBusId(bdf.encode())
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, args=(nv, h), daemon=True)
            self.thread.start()
        except Exception:
            self.thread = None
            try:
                q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
                p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                     stdout=f, stderr=subprocess.DEVNULL)
                self.fallback = (p, f)
            except Exception:
                self.fallback = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": "nvml, 20 ms poll, samples inside the timed region"}
        if self.thread is not None:
            time.sleep(0.05)
            self.stop_flag.set()
            self.thread.join(timeout=2)
            inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= (self.t1 or 1e300)]
            use = inside or self.samples[-5:]
            if not inside:
                out["source"] = "nvml, 20 ms poll, last samples before the region ended (region shorter than the poll)"
            if use:
                sm = sorted(s[1] for s in use)
                mask = 0
                for s in use:
                    mask |= s[2]
                out.update(sm_mhz=sm[len(sm) // 2], reasons=sorted(v for k, v in self.REASONS.items() if mask & k),
                           samples=len(use), power_w_max=max(s[3] for s in use))
            return out
        if self.fallback is not None:
            p, f = self.fallback
            time.sleep(0.12)
            p.terminate()
            try:
                p.wait(timeout=5)
            except Exception:
                p.kill()
            f.flush(); f.seek(0)
            sm, mx, reasons, power = [], [], set(), []
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in f.read().splitlines():
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
                os.unlink(f.name)
            except OSError:
                pass
            if sm:
                sm.sort()
                out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                           power_w_max=max(power) if power else None, source="nvidia-smi -lms 50 (whole run)")
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
        n = synth.lengths_normal(rng, S, 100.0, 15.0, 3, 200)
    elif law == "loguniform":
        n = synth.lengths_log_uniform(rng, S)
    elif law == "heavy":
        n = synth.lengths_heavy_tail(rng, S)
    elif law == "uniform40":
        n = synth.lengths_uniform(rng, S, 40, 139)
    else:
        n = synth.lengths_uniform(rng, S, 20, 119)
    return synth.random_walk_csr(n, seed)


def run_reference(args):
    """--impl reference: the CPU path on all host cores, same metric/config; each step = one bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = args.cfg
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    per_step = min(max(cores * 250, 2000), cfg["S"])     # ~0.3 ms per polyline per core -> ~0.1-0.3 s per step
    pts, off = host_sample(per_step, cfg["law"], cfg["seed"])
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
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "sample": f"{per_step} polylines ({P} points) per step"},
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
def time_steps(step, steps, warmup, stream, dist, dev, on_begin=None, on_end=None):
    """W untimed + K timed calls of step(events) on `stream`; returns (ms_total max over ranks, [per-step metrics-call ms])."""
    import torch
    for _ in range(max(warmup, 3)):
        step(None)
    torch.cuda.synchronize(dev)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if on_begin:
        on_begin()
    e0.record(stream)
    for i in range(steps):
        step(kev[i])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    if on_end:
        on_end()
    if dist is not None:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    k_ms = [a.elapsed_time(b) for a, b in kev]
    if dist is not None:
        t = torch.tensor([ms_total, sum(k_ms) / len(k_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), k_ms, float(t[1])
    return ms_total, k_ms, sum(k_ms) / len(k_ms)


def e2e_file_leg(ctx):
    """file -> (df_sl, df_bundle): BASELINE configs[0] through compute_streamline_metrics(vtk_path, max_streamlines=1000)
    and configs[1] (16 tracts x 4 timepoints of ~5k polylines) through the batched driver's loader + one device call.
    Files are synthetic legacy VTK (binary, `POINTS n double`), written to a temporary directory outside the timed region."""
    import numpy as np
    from lesion_condition_vae_b200 import synth, tract_driver as td, tract_geom_proc as tgp, vtk_io
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        pts, off = synth.config1(S=1000, seed=0)
        path = vtk_io.write_polylines(os.path.join(tmp, "cfg0.vtk"), pts, off, binary=True, point_dtype="double")
        for _ in range(3):                                               # warm-up: the pinned arena settles on one block (pinned
            tgp.compute_streamline_metrics(path, max_streamlines=1000)    # allocations are slow in a process that already pins 24 GB)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            df_sl, df_b = tgp.compute_streamline_metrics(path, max_streamlines=1000)
        dt = (time.perf_counter() - t0) / reps
        res["config0"] = {"value": len(df_sl) / dt, "unit": UNIT, "ms_per_call": 1e3 * dt, "rows": int(len(df_sl)), "file_bytes": os.path.getsize(path),
                          "call": "compute_streamline_metrics(vtk_path, max_streamlines=1000): parse + H2D + kernels + D2H + two DataFrames"}
        files = []
        for t in range(16):
            for tp in range(4):
                p, o = synth.config2_bundle(t, tp, S=5000)
                files.append(vtk_io.write_polylines(os.path.join(tmp, f"b{t}_{tp}.vtk"), p, o, binary=True, point_dtype="double"))
        nbytes = sum(os.path.getsize(f) for f in files)
        for _ in range(2):                                               # warm-up (arena and device scratch reach their final size)
            td.compute_files(files, ctx=ctx)
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            n_sl, means = td.compute_files(files, ctx=ctx)
        dt = (time.perf_counter() - t0) / reps
        res["config1"] = {"value": float(np.sum(n_sl)) / dt, "unit": UNIT, "ms_per_call": 1e3 * dt, "rows": int(np.sum(n_sl)), "bundles": len(files),
                          "file_bytes": nbytes, "call": "tract_driver.compute_files(64 files): parse into pinned buffers overlapped with H2D, one batched device call, 64 bundle rows"}
    return res


def run_ours(args):
    import numpy as np
    import torch
    from lesion_condition_vae_b200 import _lib, build as tgbuild, sharding, synth

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

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()                        # polls from here on: > 1 s of set-up precedes the timed region
    cfg = args.cfg
    ctx = _lib.Context(local)
    S, law = cfg["S"], cfg["law"]
    strong_main = args.scaling == "strong" and world > 1
    stream = torch.cuda.Stream(dev)           # a real (non-default) stream: kernels AND timing events live on it
    torch.cuda.set_stream(stream)
    sp = stream.cuda_stream
    assert sp != 0

    def make_tractogram(seed):
        n = synth.torch_lengths(law, S, seed, dev)
        pts, off = synth.torch_random_walk_csr(n, seed, dev)
        del n
        B = cfg["bundles"]
        bo = (np.arange(B + 1, dtype=np.int64) * (S // B)) if B > 1 else np.array([0, S], dtype=np.int64)
        bo[-1] = S
        return pts, off, bo

    # ---- weak: every rank owns its own tractogram -------------------------------------------------------------
    pts, off, bo = make_tractogram(cfg["seed"] + 1000 * rank)
    P = int(pts.shape[0])
    own = sharding.DeviceShard(ctx, pts, off, 0, S, bo)      # (copies into an array with two points of slack behind the last polyline)
    pts = own.points[:P]
    torch.cuda.empty_cache()

    def weak_step(ev):
        if ev is not None:
            ev[0].record(stream)
        ctx.metrics_dev(own.points.data_ptr(), _lib.F64, own.offsets.data_ptr(), S, own.P, own.out.data_ptr(), own.keep.data_ptr(), sp)
        if ev is not None:
            ev[1].record(stream)
        ctx.bundle_partials_dev(own.out.data_ptr(), own.keep.data_ptr(), 0, S, own.bundle_offsets, own.partial.data_ptr(), sp)
        return sharding.allgather_partials(own.partial) if world > 1 else None   # (world, B, 27): summed in rank order by the caller

    launches0 = ctx.launches
    ms_total, k_ms, k_avg_max = time_steps(weak_step, args.steps, args.warmup, stream, dist, dev,
                                           on_begin=clocks.mark_begin if rank == 0 and not strong_main else None,
                                           on_end=clocks.mark_end if rank == 0 and not strong_main else None)
    launches = (ctx.launches - launches0) * args.steps // (args.steps + max(args.warmup, 3))
    k_ms.sort()
    tot_S, tot_P = float(S), float(P)
    if world > 1:
        c = torch.tensor([tot_S, tot_P, float(launches)], dtype=torch.float64, device=dev)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        tot_S, tot_P, launches = float(c[0]), float(c[1]), int(c[2])
    # sanity: the result is real (every polyline kept, finite lengths); at N > 1 from the gathered partials
    if world > 1:
        g_sums, g_counts = sharding.combine_partials(weak_step(None).cpu().numpy())
    else:
        torch.cuda.synchronize(dev)
        g_sums, g_counts = sharding.combine_partials(own.partial.cpu().numpy()[None])
    n_kept = int(g_counts[:, 0].sum())
    mean_len = float(g_sums[:, 0].sum()) / max(int(g_counts[:, 1].sum()), 1)

    # ---- strong: ONE tractogram (same seed on every rank) cut into CSR ranges ---------------------------------
    strong = None
    if world > 1 and not args.no_strong:
        g_pts, g_off, g_bo = make_tractogram(cfg["seed"])
        bounds = sharding.shard_ranges(g_off.cpu().numpy(), world)
        shard = sharding.DeviceShard(ctx, g_pts, g_off, bounds[rank], bounds[rank + 1], g_bo, copy=True)   # own copy: the global table is freed
        g_P = int(g_pts.shape[0])
        del g_pts, g_off
        torch.cuda.empty_cache()

        def strong_step(ev):
            if ev is not None:
                ev[0].record(stream)
            shard.compute(sp)
            if ev is not None:
                ev[1].record(stream)
            return sharding.allgather_partials(shard.partial)

        s_ms, s_k, s_kmax = time_steps(strong_step, args.steps, args.warmup, stream, dist, dev,
                                       on_begin=clocks.mark_begin if rank == 0 and strong_main else None,
                                       on_end=clocks.mark_end if rank == 0 and strong_main else None)
        s_sums, s_counts = sharding.combine_partials(strong_step(None).cpu().numpy())
        sh = torch.tensor([float(shard.S), float(shard.P)], dtype=torch.float64, device=dev)
        allsh = [torch.zeros_like(sh) for _ in range(world)]
        dist.all_gather(allsh, sh)
        strong = {"metric": METRIC, "value": S * args.steps / (s_ms / 1e3), "unit": UNIT, "scaling": "strong",
                  "ms_per_step": s_ms / args.steps, "streamlines_total": S, "points_total": g_P,
                  "points_per_sec": g_P * args.steps / (s_ms / 1e3),
                  "shard_polylines": [int(x[0]) for x in allsh], "shard_points": [int(x[1]) for x in allsh],
                  "kernels_ms_max_rank": s_kmax,
                  "check": {"n_streamlines": int(s_counts[:, 0].sum()), "length_mean": float(s_sums[:, 0].sum()) / max(int(s_counts[:, 1].sum()), 1)},
                  "note": "one tractogram (seed of the config, identical on every rank) cut by sharding.shard_ranges; per step: "
                          "tg_metrics_csr_dev + tg_bundle_partials_dev on the slice, one NCCL all-gather of B x 27 doubles; "
                          "efficiency vs N = 1 is value(N) / (N * value(1)) with value(1) = the N = 1 line's `value`"}
        del shard
        torch.cuda.empty_cache()

    # ---- e2e: HOST buffers through tg_metrics_csr_host (pinned), H2D + kernels + D2H per step ----
    want = args.e2e_streamlines or (S if world == 1 else min(S, 2_000_000))
    Se = min(want, S)
    need = 24 * int(off[Se].item()) * 1.6 + 200 * Se
    if need * max(world, 1) > 0.5 * mem_available_bytes():          # never drive the host out of memory
        Se = min(S, 1_000_000)
    Pe = int(off[Se].item())
    h_pts = torch.empty((Pe, 3), dtype=torch.float64, pin_memory=True)
    h_off = torch.empty(Se + 1, dtype=torch.int64, pin_memory=True)
    h_pts.copy_(pts[:Pe]); h_off.copy_(off[:Se + 1])
    torch.cuda.synchronize(dev)
    hp, ho = h_pts.numpy(), h_off.numpy()
    h_out = torch.empty((17, Se), dtype=torch.float64, pin_memory=True).numpy()      # result buffers the caller owns, pinned
    h_keep = torch.empty(Se, dtype=torch.uint8, pin_memory=True).numpy()
    bo_e = bo if Se == S else np.array([0, Se], dtype=np.int64)
    e2e_steps = max(3, min(args.steps, 10)) if Se <= 2_000_000 else 3
    ctx.metrics_host(hp, ho, bo_e, out=h_out, keep=h_keep)             # warm-up (allocates device scratch)
    ctx.metrics_host(hp, ho, bo_e, out=h_out, keep=h_keep)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        o_h, k_h, s_h, c_h = ctx.metrics_host(hp, ho, bo_e, out=h_out, keep=h_keep)
    e2e_sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t[0])
    e2e_value = world * Se * e2e_steps / e2e_sec
    nB = len(bo_e) - 1
    h2d = 24 * Pe + 8 * (Se + 1)
    d2h = 17 * 8 * Se + Se + nB * (13 * 8 + 14 * 8)
    assert args.no_check or (int(c_h[:, 0].sum()) == Se and np.isfinite(o_h[0]).all())
    # same call with the points stored as float32 on the host (what legacy-VTK tract files hold; upcast exactly on
    # the device): half the H2D bytes.  Reported beside `e2e`, never instead of it.
    S32 = min(Se, 1_000_000)
    P32 = int(ho[S32])
    h_pts32 = torch.empty((P32, 3), dtype=torch.float32, pin_memory=True)
    h_pts32.copy_(pts[:P32])
    torch.cuda.synchronize(dev)
    hp32, ho32 = h_pts32.numpy(), ho[:S32 + 1].copy()
    ctx.metrics_host(hp32, ho32)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.metrics_host(hp32, ho32)
    e2e32_sec = (time.perf_counter() - t0) / 3
    if world > 1:
        t = torch.tensor([e2e32_sec], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e32_sec = float(t[0])
    del h_pts32, hp32

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    clk = clocks.stop()
    peak, peak_src = measured_peaks()
    abytes = algorithmic_bytes(P, S)
    achieved = abytes / (k_avg_max * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            if tj.get("streamlines") == S and tj.get("law") == law:
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = f"not measured in this run: {tj.get('source')} (profiles/traffic.json, build {tj.get('build_id', 'round 1')})"
        except Exception:
            pass
    sass = tgbuild.steady_block_stats(_lib.LIB_PATH)
    ops = sass["fp64_pipe_per_point"] if sass else None
    pipe_peak = 60 * 148 * 1.965e9                      # lane-operations / s at the maximum clock (tools/microbench.cu: 60 / clk / SM)
    ms_step = ms_total / args.steps
    sec = ms_total / 1e3
    weak_value = tot_S * args.steps / sec
    line = {
        "metric": METRIC, "value": weak_value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "build_id": _lib.build_id(),
        "config": {"workload": workload_name(cfg), "streamlines_per_gpu": S, "points_per_gpu": P, "bundles": cfg["bundles"],
                   "parallelism": f"csr-range shards x{world}" + (", NCCL all-gather of 27 bundle partials/rank" if world > 1 else "")
                   + (f", ranks bound to their GPU's NUMA node (rank 0: node {numa})" if numa is not None else ""),
                   "l2": "inputs (24 B/point, >> 126 MB L2) stream from HBM every step; no flush needed" if 24 * P > 4 * 126e6 else
                         "inputs fit in the 126 MB L2: this small configuration is launch-latency bound, its roofline fraction is not a bandwidth statement",
                   "e2e_workload": (f"the whole workload: {Se} polylines ({Pe} points)" if Se == S else f"first {Se} polylines ({Pe} points)")
                                   + " per GPU from pinned host buffers, full df_sl table copied back"},
        "points_per_sec": tot_P * args.steps / sec,
        "kernel_ms": {"metrics_avg": k_avg_max, "metrics_min": k_ms[0], "metrics_median": k_ms[len(k_ms) // 2]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes,
                     "kernel": "tg_metrics_csr_dev (queue kernels + k_metrics_grouped + k_metrics_long: 17 metrics per polyline), avg CUDA-event time over the timed steps",
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        # the unit that actually binds the kernel: fp64-pipe instructions per interior point, counted in the
        # SASS of the loaded library, against 60 lane-operations / clock / SM (tools/microbench.cu)
        "fp64_pipe": None if ops is None else {
            "lane_ops_per_point": ops, "instructions_per_point": sass["instructions_per_point"], "source": sass["source"],
            "achieved_tlaneops": ops * P / (k_avg_max * 1e-3) / 1e12, "peak_tlaneops_at_max_clock": pipe_peak / 1e12,
            "frac": ops * P / (k_avg_max * 1e-3) / pipe_peak},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_sec / e2e_steps, "steps": e2e_steps},
        "e2e_f32_points": {"value": world * S32 / e2e32_sec, "unit": UNIT, "h2d_bytes_per_step": 12 * P32 + 8 * (S32 + 1),
                           "ms_per_step": 1e3 * e2e32_sec, "streamlines": S32,
                           "note": "same call, host points stored as float32 (legacy-VTK float files), computed in fp64"},
        "gpu_launches": int(launches),
        "clocks": clk,
        "check": {"n_streamlines": n_kept, "length_mean": mean_len},
    }
    if strong is not None:
        line["strong"] = strong
        if strong_main:                                  # --scaling strong: the sharded tractogram is the headline
            line.update(value=strong["value"], ms_per_step=strong["ms_per_step"], scaling="strong", points_per_sec=strong["points_per_sec"])
            line["weak"] = {"value": weak_value, "ms_per_step": ms_step}
            line["config"]["parallelism"] = f"ONE tractogram of {S} polylines cut into {world} CSR ranges (balanced by points), NCCL all-gather of 27 bundle partials/rank"
    if world == 1 and not args.no_e2e_file:
        try:
            line["e2e_file"] = e2e_file_leg(ctx)
        except Exception as e:                           # never lose the line over the file leg
            line["e2e_file"] = {"error": f"{type(e).__name__}: {e}"}
    if not args.no_cpu_baseline and world == 1:
        from oracle import cpu_bench
        cs = min(args.cpu_sample, Se)
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        cp = hp[:int(ho[cs])].copy(); co = ho[:cs + 1].copy()
        rows, dt = cpu_bench.timed_run(cp, co, cores)
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
