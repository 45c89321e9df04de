import os, sys, tempfile, time
sys.path.insert(0, "/root/repo")
import numpy as np
from lesion_condition_vae_b200 import _lib, synth, tract_geom_proc as tgp, vtk_io
with tempfile.TemporaryDirectory() as tmp:
    pts, off = synth.config1(S=1000, seed=0)
    path = vtk_io.write_polylines(os.path.join(tmp, "cfg0.vtk"), pts, off, binary=True, point_dtype="double")
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(20): tgp.compute_streamline_metrics(path, max_streamlines=1000)
        print("full call ms", (time.perf_counter() - t0) / 20 * 1e3)
    t0 = time.perf_counter()
    for _ in range(20): p, o = tgp._load_for_device(path)
    print("load ms", (time.perf_counter() - t0) / 20 * 1e3, p.dtype)
    ctx = _lib.default_context()
    t0 = time.perf_counter()
    for _ in range(20): r = ctx.metrics_host(p, o)
    print("metrics_host (BE pinned) ms", (time.perf_counter() - t0) / 20 * 1e3)
    pn = np.ascontiguousarray(p.astype(np.float64))
    t0 = time.perf_counter()
    for _ in range(20): r = ctx.metrics_host(pn, o)
    print("metrics_host (native pageable) ms", (time.perf_counter() - t0) / 20 * 1e3)
    t0 = time.perf_counter()
    for _ in range(20): tgp.frames_from_table(r[0], r[1] == 3, r[2][0], r[3][0])
    print("frames ms", (time.perf_counter() - t0) / 20 * 1e3)
