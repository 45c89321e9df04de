"""Per-call latency of the host-buffer entry point on pageable numpy arrays (what the per-file drop-in passes).
B200, round 1: 1,000 polylines 0.48 ms, 5,000 polylines 0.97 ms, 50,000 polylines 8.3 ms (pageable H2D ~10 GB/s;
the pinned buffers of bench.py reach 53 GB/s).  Reading and parsing the same 5,000-polyline file takes ~20 ms."""
import time, numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lesion_condition_vae_b200 import _lib, synth
ctx = _lib.Context(0)
for S in (1000, 5000, 50000):
    pts, off = synth.config1(S=S, seed=1)
    ctx.metrics_host(pts, off)
    t = time.perf_counter()
    for _ in range(50): ctx.metrics_host(pts, off)
    dt = (time.perf_counter() - t) / 50
    print(S, len(pts), f"{dt*1e3:.3f} ms per call, {S/dt:.3e} polylines/s")
