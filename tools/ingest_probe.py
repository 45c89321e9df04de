#!/usr/bin/env python3
"""Where the file -> bundle-rows time of BASELINE configs[1] goes (64 binary double files x 5,000 polylines, 804 MB):
parse only (parser threads, pinned arena, no device), then the full tract_driver.compute_files.  Run on the GPU box."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lesion_condition_vae_b200 import _lib, synth, tract_driver as td, vtk_io

with tempfile.TemporaryDirectory() as tmp:
    files = []
    for t in range(16):
        for tp in range(4):
            p, o = synth.config2_bundle(t, tp, S=5000)
            files.append(vtk_io.write_polylines(os.path.join(tmp, f"b{t}_{tp}.vtk"), p, o, binary=True, point_dtype="double"))
    nbytes = sum(os.path.getsize(f) for f in files)
    ctx = _lib.default_context()
    arena = _lib.default_arena()
    import concurrent.futures as cf
    for threads in (1, 4, 8, 16):
        arena.reset()
        with cf.ThreadPoolExecutor(threads) as pool:
            list(pool.map(lambda f: td._load(f, None, arena), files))
            arena.reset()
            t0 = time.perf_counter()
            list(pool.map(lambda f: td._load(f, None, arena), files))
            dt = time.perf_counter() - t0
        print(f"parse only, {threads:2d} threads: {1e3 * dt:7.1f} ms  {nbytes / dt / 1e9:5.2f} GB/s")
    for threads in (1, 8, 16):
        td.PARSER_THREADS = threads
        td.compute_files(files, ctx=ctx)
        td.compute_files(files, ctx=ctx)
        t0 = time.perf_counter()
        for _ in range(3):
            n_sl, means = td.compute_files(files, ctx=ctx)
        dt = (time.perf_counter() - t0) / 3
        print(f"compute_files, {threads:2d} parser threads: {1e3 * dt:7.1f} ms  {nbytes / dt / 1e9:5.2f} GB/s  {n_sl.sum() / dt:.3e} polylines/s")
