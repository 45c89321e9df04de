"""TEST INFRASTRUCTURE — times the CPU oracle port on the host cores (bench.py's cpu_baseline leg
and `bench.py --impl reference`).  Not a product path; nothing under lesion_condition_vae_b200/ imports it.

The reference (/root/reference/src/geometry/tract_geom_proc.py:153-212) is a single-threaded
Python loop; to give the CPU "all the host threads it can use" the polylines are split into
contiguous ranges, one worker process per core, each running oracle.streamline_oracle on its range
and returning its 13 partial bundle sums — the same sharding the GPU path uses across devices.
Workers are spawned (not forked: the parent may hold a CUDA context) and import numpy/pandas only.
"""
from __future__ import annotations

import multiprocessing as mp
import time

import numpy as np

_SRC = (0, 2, 4, 6, 7, 8, 10, 11, 12, 16, 13, 14, 15)   # df_sl column feeding bundle column j (ref:197-209)


def _work(args):
    pts, off = args
    from oracle import streamline_oracle as so
    table, _ = so.per_streamline_table(pts, off)
    if len(table) == 0:
        return 0, np.zeros(13), np.zeros(13, np.int64)
    cols = table[:, _SRC]
    ok = ~np.isnan(cols)
    return len(table), np.where(ok, cols, 0.0).sum(axis=0), ok.sum(axis=0)


def _split(pts, off, parts):
    S = len(off) - 1
    tgt = np.linspace(0, int(off[-1]), parts + 1)
    cuts = np.unique(np.clip(np.searchsorted(off, tgt), 0, S))
    if cuts[0] != 0:
        cuts = np.concatenate([[0], cuts])
    if cuts[-1] != S:
        cuts = np.concatenate([cuts, [S]])
    jobs = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        if b > a:
            jobs.append((pts[off[a]:off[b]], off[a:b + 1] - off[a]))
    return jobs


class Pool:
    def __init__(self, cores):
        self.cores = max(1, int(cores))
        self.pool = mp.get_context("spawn").Pool(self.cores) if self.cores > 1 else None
        if self.pool is not None:                       # make every worker import numpy/pandas/oracle now
            tiny = (np.zeros((3, 3)) + np.arange(3)[:, None], np.array([0, 3], np.int64))
            self.pool.map(_work, [tiny] * (2 * self.cores))

    def run(self, pts, off):
        """Rows produced (after both filters) for the tractogram, using every worker."""
        jobs = _split(pts, off, self.cores * 4)
        res = self.pool.map(_work, jobs, chunksize=1) if self.pool is not None else [_work(j) for j in jobs]
        return int(sum(r[0] for r in res))

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool.join()


def timed_run(pts, off, cores):
    """(rows, seconds) for one pass over (pts, off) on `cores` worker processes (pool start-up untimed)."""
    pool = Pool(cores)
    try:
        t0 = time.perf_counter()
        rows = pool.run(pts, off)
        return rows, time.perf_counter() - t0
    finally:
        pool.close()
