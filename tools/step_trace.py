#!/usr/bin/env python3
"""Per-step time of tg_metrics_csr_dev over a long run, with the SM clock / power / clock-event reasons NVML reports between
steps: how the 10M-polyline step goes from ~8.4 ms (first steps) to ~9.2 ms (sustained).  usage: python tools/step_trace.py [steps] [gap_ms]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, pynvml as nv
from lesion_condition_vae_b200 import _lib, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
gap = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
dev = torch.device("cuda:0")
S = 10_000_000
n = synth.torch_lengths("normal", S, 5, dev)
pts, off = synth.torch_random_walk_csr(n, 5, dev)
P = pts.shape[0]
out = torch.empty((17, S), dtype=torch.float64, device=dev); keep = torch.empty(S, dtype=torch.uint8, device=dev)
ctx = _lib.Context(0)
nv.nvmlInit(); h = nv.nvmlDeviceGetHandleByIndex(0)
st = torch.cuda.Stream(dev)
torch.cuda.synchronize()
time.sleep(2.0)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
info = []
for i in range(steps):
    ev[i][0].record(st)
    ctx.metrics_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, out.data_ptr(), keep.data_ptr(), st.cuda_stream)
    ev[i][1].record(st)
    if i % 4 == 3 or gap > 0:
        st.synchronize()
        info.append((i, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_MEM), nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                     nv.nvmlDeviceGetCurrentClocksEventReasons(h), nv.nvmlDeviceGetTemperature(h, 0)))
        if gap > 0:
            time.sleep(gap / 1e3)
torch.cuda.synchronize()
ms = [a.elapsed_time(b) for a, b in ev]
print("ms:", " ".join(f"{m:.2f}" for m in ms))
for r in info:
    print("step %3d  sm %4d MHz  mem %4d MHz  %6.1f W  reasons 0x%x  %d C" % r)
print("limits: power limit %.0f W (enforced %.0f W), max sm %d" % (nv.nvmlDeviceGetPowerManagementLimit(h) / 1000.0, nv.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0,
                                                                     nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
