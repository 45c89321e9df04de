#!/usr/bin/env python
"""Device-resident timing of the arc-length resampling kernel on the bench.py workload
shape: S polylines x ~100 points -> 100 nodes each.  Prints one JSON line with the HBM roofline of the kernel:
algorithmic bytes = 24 P + 8 (S+1) read, 24 K S written.  Not the headline bench (bench.py is)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streamlines", type=int, default=10_000_000)
    ap.add_argument("--nodes", type=int, default=100)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--law", default="normal")
    args = ap.parse_args()
    import torch
    from lesion_condition_vae_b200 import _lib, synth
    dev = torch.device("cuda:0")
    S, K = args.streamlines, args.nodes
    n = synth.torch_lengths(args.law, S, 5, dev)
    pts, off = synth.torch_random_walk_csr(n, 5, dev)
    P = pts.shape[0]
    nodes = torch.empty((S, K, 3), dtype=torch.float64, device=dev)
    ctx = _lib.Context(0)
    stream = torch.cuda.Stream(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            ctx.resample_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, K, nodes.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize(dev)
        for a, b in ev:
            a.record(stream)
            ctx.resample_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, K, nodes.data_ptr(), stream.cuda_stream)
            b.record(stream)
        torch.cuda.synchronize(dev)
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    avg = sum(ms) / len(ms)
    abytes = 24 * P + 8 * (S + 1) + 24 * K * S
    peak = 6555.2
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    achieved = abytes / (avg * 1e-3) / 1e9
    print(json.dumps({"kernel": "k_resample", "streamlines": S, "points": P, "nodes": K, "ms_avg": avg, "ms_min": ms[0],
                      "streamlines_per_sec": S / (avg * 1e-3),
                      "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                   "algorithmic_bytes_per_launch": abytes}}))


if __name__ == "__main__":
    main()
