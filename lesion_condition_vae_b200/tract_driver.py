"""Batched tract-geometry driver: the B200-first form of the reference's
``src/geometry/comprehensive_tract_geometry_analysis.py``.

The reference walks groups x subjects x timepoints x 16 tracts serially
(comprehensive_tract_geometry_analysis.py:169-195), gunzips every ``.vtk.gz`` to a sibling file
(:54-76) and calls ``compute_streamline_metrics(path, max_streamlines)`` once per file (:102) —
2,368 calls per run.  Here the files of a batch (default: everything) are parsed into ONE CSR
tractogram with a bundle table, the ``max_streamlines`` prefix rule is resolved on the host, and
the metrics + bundle reduction run as ONE device call (``compute_bundles`` hook); ``.gz`` is read
from memory.  The result is the same DataFrame / CSV: one row per tract file that produced at
least one streamline, 14 bundle columns in the reference order followed by ``subject_id,
timepoint, tract, group`` (:109-115), rows in the reference's loop order, ``n_streamlines`` as
float (the reference passes the summary through ``Series.to_dict``, :109).

Same public names as the reference module: ``TRACT_LIST``, ``load_config``, ``get_all_subjects``,
``process_single_tract``, ``process_all_tracts``, ``generate_summary_statistics``, ``main``.
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Callable, Optional

import numpy as np
import pandas as pd

from . import vtk_io
from .tract_geom_proc import BUNDLE_COLUMNS

TRACT_LIST = [  # comprehensive_tract_geometry_analysis.py:25-32
    'chip_right', 'hipcom', 'thalsub_left',
    'cing_left', 'thalsub_right',
    'cing_right',
    'fimbria_left', 'ant_comm', 'fimbria_right',
    'atr_left', 'fornix_left', 'intcap_left',
    'atr_right', 'chip_left', 'fornix_right', 'intcap_right'
]
TIMEPOINTS = ['2d', '9d', '1mo', '5mo']  # :162
META_COLUMNS = ('subject_id', 'timepoint', 'tract', 'group')  # :112-115


def load_config(config_path=None):
    """Tract configuration with subject metadata.  The reference looks for
    ``<module dir>/lesion_vae_analysis/configs/tract_config.json`` (:36), a path that does not exist
    in its own tree; here the path is an argument, defaulting to ``configs/tract_config.json``
    two levels above the package when present."""
    if config_path is None:
        for cand in (Path.cwd() / "configs" / "tract_config.json",
                     Path(__file__).resolve().parent.parent / "configs" / "tract_config.json"):
            if cand.exists():
                config_path = cand
                break
        else:
            raise FileNotFoundError("tract_config.json not found; pass config_path")
    with open(config_path, 'r') as f:
        return json.load(f)


def get_all_subjects(config):
    """:41-51 — {group: [subject id strings]} for Sham, TBI, PTE, in config order."""
    out = {}
    for group, subject_list in config.get('groups', {}).items():
        if group in ['Sham', 'TBI', 'PTE']:
            out[group] = [str(s) for s in subject_list]
    return out


def find_tract_file(data_dir, subject_id, timepoint, tract_name):
    """:86-93 — ``<data>/<subject>/<timepoint>/bundles/<tract>_curves.vtk.gz``, else without ``.gz``."""
    base = Path(data_dir) / subject_id / timepoint / "bundles"
    for name in (f"{tract_name}_curves.vtk.gz", f"{tract_name}_curves.vtk"):
        p = base / name
        if p.exists():
            return p
    return None


def select_prefix(points, offsets, max_streamlines):
    """The loader rule of tract_geom_proc.py:17-25 on a CSR tractogram, vectorised: indices of the
    polylines with more than 2 points and only finite coordinates, in file order, cut after
    ``max_streamlines`` of them (the cap is tested after an append, so a cap <= 0 still admits one).

    With a cap only the candidates actually needed are inspected (a file of 5,000 polylines capped at 100 costs
    100 polylines' worth of finite tests); points may be in the file's big-endian storage."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = np.diff(offsets)
    if len(n) == 0:
        return np.zeros(0, dtype=np.int64)
    cand = np.flatnonzero(n > 2)
    if max_streamlines is None:
        want = len(cand)
    else:
        want = max(int(max_streamlines), 1)
    keep = []
    have, pos = 0, 0
    while have < want and pos < len(cand):
        take = cand[pos:pos + max(want - have, 64)]
        pos += len(take)
        lo, hi = int(offsets[take[0]]), int(offsets[take[-1] + 1])
        block = np.asarray(points[lo:hi])
        bad_pt = ~np.isfinite(block.astype(block.dtype.newbyteorder("="), copy=False)).all(axis=1)
        csum = np.concatenate([[0], np.cumsum(bad_pt, dtype=np.int64)])
        ok = (csum[offsets[take + 1] - lo] - csum[offsets[take] - lo]) == 0
        good = take[ok][:want - have]
        keep.append(good)
        have += len(good)
    return np.concatenate(keep) if keep else np.zeros(0, dtype=np.int64)


def gather_polylines(points, offsets, idx):
    """CSR of the selected polylines (contiguous copy)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = offsets[idx + 1] - offsets[idx]
    new_off = np.zeros(len(idx) + 1, dtype=np.int64)
    np.cumsum(n, out=new_off[1:])
    if len(idx) and np.array_equal(idx, np.arange(idx[0], idx[0] + len(idx))):
        return points[offsets[idx[0]]:offsets[idx[-1] + 1]], new_off          # a contiguous run: a view
    rows = np.repeat(offsets[idx] - new_off[:-1], n) + np.arange(int(new_off[-1]), dtype=np.int64)
    return points[rows], new_off


def _default_compute(points, offsets, bundle_offsets):
    """One device call for an in-memory batch -> (n_streamlines int64[B], means float64[B,13])."""
    from . import _lib
    ctx = _lib.default_context()
    _, _, sums, counts = ctx.metrics_host(points, offsets, bundle_offsets, want_rows=False)
    return counts[:, 0].copy(), _means(sums, counts)


def _means(sums, counts):
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.where(counts[:, 1:] > 0, sums / np.maximum(counts[:, 1:], 1), np.nan)


_DEVICE_COMPUTE = _default_compute          # process_batch streams files to the device unless this module attribute was replaced
PARSER_THREADS = 8                # files parsed concurrently ahead of the push (profiles/r2_ingest_probe.txt: 1 -> 127 ms, 8 -> 26 ms per 64 files)
MAX_BATCH_POINTS = 1 << 27        # a device call is flushed beyond this many points (3.2 GB of float64 coordinates)


def _load(path, max_streamlines, arena):
    """One tract file -> (points, local offsets) of the polylines the reference's loader would hand over, or None.

    ``max_streamlines=None``: the whole file as it is (the loader filter is evaluated on the device, `keep` flags);
    otherwise the prefix rule (select_prefix).  Points stay in the file's storage type and — with an arena — in
    pinned memory."""
    pts, off = vtk_io.read_polylines_raw(path, arena)
    if max_streamlines is None:
        return (pts, off) if len(off) > 1 else None
    idx = select_prefix(pts, off, max_streamlines)
    if len(idx) == 0:
        return None
    return gather_polylines(pts, off, idx)


def compute_files(paths, max_streamlines=None, ctx=None, arena=None, on_error=None):
    """The files of a batch through ONE device call: -> (n_streamlines int64[F], means float64[F,13]), one bundle
    per file; an unreadable or empty file gives n_streamlines 0 and NaN means (``on_error(i, exc)`` is told).

    Pipeline (comprehensive_tract_geometry_analysis.py:169-195 is the serial loop replaced): file i
    is parsed into pinned memory and PUSHED (tg_batch_push queues its host-to-device copy and returns), so parsing
    file i+1 overlaps the transfer and decode of file i; tg_batch_run then computes every polyline and reduces the
    bundles.  Batches are flushed at MAX_BATCH_POINTS points.  A device failure of a batch falls back to one call per
    file, so that only the failing tracts are lost (the reference isolates failures per tract, :95-131)."""
    from . import _lib
    ctx = ctx or _lib.default_context()
    if arena is None:
        try:
            arena = _lib.default_arena()
        except _lib.TractGeomError:
            arena = None                                   # no pinned memory: pageable copies, same results
    F = len(paths)
    n_sl = np.zeros(F, dtype=np.int64)
    means = np.full((F, _lib.N_BUNDLE_COLS), np.nan)

    def run(members, bo):
        _, _, sums, counts = ctx.batch_run(np.asarray(bo, dtype=np.int64))
        m = _means(sums, counts)
        for b, i in enumerate(members):
            n_sl[i], means[i] = counts[b, 0], m[b]

    def one_by_one(members):
        for i in members:
            try:
                item = _load(paths[i], max_streamlines, None)
                if item is None:
                    continue
                _, _, sums, counts = ctx.metrics_host(item[0], item[1], want_rows=False)
                n_sl[i], means[i] = counts[0, 0], _means(sums, counts)[0]
            except Exception as e:
                if on_error:
                    on_error(i, e)

    # parser threads run ahead of the pushing thread (file reads and the native cell walk release the GIL): the order
    # of the pushes — hence of the bundles — stays the file order
    import concurrent.futures as cf
    import os
    workers = max(1, min(PARSER_THREADS, F, os.cpu_count() or 1))
    pool = cf.ThreadPoolExecutor(workers) if workers > 1 else None
    ahead = {}

    def fetch(k):
        """(points, offsets) | None | the exception, for file k; keeps `workers` files in flight."""
        if pool is None:
            try:
                return _load(paths[k], max_streamlines, arena)
            except Exception as e:
                return e
        for j in range(k, min(k + 2 * workers, F)):
            if j not in ahead:
                ahead[j] = pool.submit(_load, paths[j], max_streamlines, arena)
        try:
            return ahead.pop(k).result()
        except Exception as e:
            return e

    i = 0
    while i < F:
        if arena is not None and not ahead:
            arena.reset()
        members, bo, pushed, held = [], [0], 0, []
        try:
            ctx.batch_begin(1 << 20, 1 << 14)
            while i < F and pushed < MAX_BATCH_POINTS:
                item = fetch(i)
                if isinstance(item, Exception):             # the reference prints and skips (:129-131)
                    if on_error:
                        on_error(i, item)
                    item = None
                if item is not None:
                    held.append(ctx.batch_push(item[0], item[1]))     # keeps the (pinned) array alive until the run
                    members.append(i)
                    bo.append(bo[-1] + len(item[1]) - 1)
                    pushed += int(item[1][-1])
                i += 1
            if members:
                run(members, bo)
        except _lib.TractGeomError as e:                    # device-side failure of the batch: isolate it per file
            if on_error:
                on_error(-1, e)
            one_by_one(members)
        del held
    if pool is not None:
        pool.shutdown(wait=True)
    return n_sl, means


def process_batch(jobs, max_streamlines=None, compute: Optional[Callable] = None, verbose=False):
    """``jobs`` = list of (subject_id, timepoint, tract_name, group, path).  Returns one metrics dict per
    job, or None where the reference would have skipped the tract (unreadable file, :129-131, or no
    surviving streamline, which raises KeyError('length') inside the reference's hot path).

    ``compute`` = a hook ``(points, offsets, bundle_offsets) -> (n_streamlines, means)`` for an in-memory batch
    (the CPU tests plug the oracle in here); by default the files stream through :func:`compute_files`."""
    results = [None] * len(jobs)
    errors = {}
    if compute is None and _default_compute is not _DEVICE_COMPUTE:
        compute = _default_compute                     # a test (or a caller) swapped the in-memory hook
    if compute is None:
        n_all, means_all = compute_files([j[4] for j in jobs], max_streamlines, on_error=lambda i, e: errors.__setitem__(i, e))
        live = list(range(len(jobs)))
        n_sl, means = n_all, means_all
    else:
        P, O, B, live = [], [np.zeros(1, np.int64)], [0], []
        base = 0
        dtype = None
        for j, (_, _, tract, _, path) in enumerate(jobs):
            try:
                pts, off = vtk_io.read_polylines_csr(path)
            except Exception as e:  # the reference prints and skips (:129-131)
                errors[j] = e
                continue
            idx = select_prefix(pts, off, max_streamlines)
            if len(idx) == 0:
                continue
            p, o = gather_polylines(pts, off, idx)
            dtype = p.dtype if dtype is None else np.result_type(dtype, p.dtype)
            P.append(p); O.append(o[1:] + base); base += int(o[-1]); B.append(B[-1] + len(idx)); live.append(j)
        if live:
            points = np.concatenate([np.asarray(p, dtype=dtype) for p in P]) if len(P) > 1 else np.ascontiguousarray(P[0], dtype=dtype)
            n_sl, means = compute(points, np.concatenate(O), np.asarray(B, dtype=np.int64))
        else:
            n_sl, means = np.zeros(0, np.int64), np.zeros((0, 13))
    for b, j in enumerate(live):
        subject_id, timepoint, tract, group, _ = jobs[j]
        if j in errors and verbose:
            print(f"      [ERROR] Failed to process {tract}: {errors[j]}")
        if n_sl[b] == 0:                               # unreadable, nothing selected, or every selected polyline had length <= 1e-8
            if verbose and j not in errors:
                print(f"      [ERROR] Failed to process {tract}: 'length'")
            continue
        m = {"n_streamlines": float(n_sl[b])}          # Series.to_dict of a mixed int/float row gives floats (:109)
        for name, v in zip(BUNDLE_COLUMNS[1:], means[b]):
            m[name] = float(v)
        m['subject_id'] = subject_id; m['timepoint'] = timepoint; m['tract'] = tract; m['group'] = group
        results[j] = m
        if verbose:
            print(f"      ✓ {tract}: {m['n_streamlines']} streamlines, length={m['length_mean']:.1f}mm")
    return results


def process_single_tract(subject_id, timepoint, tract_name, data_dir, group, max_streamlines=None, compute=None):
    """:79-131 — one tract -> metrics dict or None."""
    path = find_tract_file(Path(data_dir), subject_id, timepoint, tract_name)
    if path is None:
        return None
    return process_batch([(subject_id, timepoint, tract_name, group, path)], max_streamlines, compute)[0]


def process_all_tracts(config, data_dir, output_dir, max_streamlines=None, compute=None, batch="all", verbose=False):
    """:134-220 — every tract of every subject and timepoint -> DataFrame (reference row order).

    ``batch``: "all" (one device call for the whole study), "subject" or "timepoint" (one call per
    subject / per subject x timepoint, for studies that do not fit in host memory at once)."""
    data_dir, output_dir = Path(data_dir), Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    jobs, cuts = [], []
    for group, subjects in get_all_subjects(config).items():
        for subject_id in sorted(subjects):
            for timepoint in TIMEPOINTS:
                for tract in TRACT_LIST:
                    path = find_tract_file(data_dir, subject_id, timepoint, tract)
                    if path is not None:
                        jobs.append((subject_id, timepoint, tract, group, path))
                if batch == "timepoint":
                    cuts.append(len(jobs))
            if batch == "subject":
                cuts.append(len(jobs))
    cuts = sorted(set(cuts + [len(jobs)]))
    results, lo = [], 0
    for hi in cuts:
        if hi > lo:
            results += [m for m in process_batch(jobs[lo:hi], max_streamlines, compute, verbose) if m is not None]
        lo = hi
    return pd.DataFrame(results)


def generate_summary_statistics(results_df, output_dir):
    """:223-296 — the two summary CSVs (by group x timepoint, by tract x group)."""
    output_dir = Path(output_dir)
    key_metrics = ['length_mean', 'tortuosity_mean', 'curv_mean_avg', 'elongation_ratio_mean', 'planarity_ratio_mean']
    rows = []
    for group in sorted(results_df['group'].unique()):
        for tp in sorted(results_df['timepoint'].unique()):
            sub = results_df[(results_df['group'] == group) & (results_df['timepoint'] == tp)]
            if len(sub) > 0:
                r = {'group': group, 'timepoint': tp, 'n_records': len(sub),
                     'n_subjects': sub['subject_id'].nunique(), 'n_tracts': sub['tract'].nunique()}
                for metric in key_metrics:
                    if metric in sub.columns:
                        r[f'{metric}_mean'] = sub[metric].mean()
                        r[f'{metric}_std'] = sub[metric].std()
                rows.append(r)
    summary_df = pd.DataFrame(rows)
    summary_df.to_csv(output_dir / "summary_statistics_by_group_timepoint.csv", index=False)
    rows = []
    for tract in sorted(results_df['tract'].unique()):
        for group in sorted(results_df['group'].unique()):
            sub = results_df[(results_df['tract'] == tract) & (results_df['group'] == group)]
            if len(sub) > 0:
                rows.append({'tract': tract, 'group': group, 'n_records': len(sub),
                             'length_mean': sub['length_mean'].mean(), 'length_std': sub['length_mean'].std(),
                             'tortuosity_mean': sub['tortuosity_mean'].mean(), 'tortuosity_std': sub['tortuosity_mean'].std(),
                             'curv_mean': sub['curv_mean_avg'].mean(), 'curv_std': sub['curv_mean_avg'].std()})
    tract_summary_df = pd.DataFrame(rows)
    tract_summary_df.to_csv(output_dir / "summary_statistics_by_tract_group.csv", index=False)
    return summary_df, tract_summary_df


def main(data_dir=None, output_dir=None, config_path=None, max_streamlines=100):
    """:299-329 — run the study and write ``comprehensive_tract_geometry_metrics.csv`` plus the two
    summaries.  ``max_streamlines=100`` is the reference's shipped setting (:310); None = all."""
    root = Path.cwd()
    data_dir = Path(data_dir) if data_dir else root / "data"
    output_dir = Path(output_dir) if output_dir else root / "results" / "comprehensive_tract_geometry"
    config = load_config(config_path)
    results_df = process_all_tracts(config, data_dir, output_dir, max_streamlines=max_streamlines)
    if len(results_df) == 0:
        print("[ERROR] No results to save!")
        return results_df
    results_df.to_csv(output_dir / "comprehensive_tract_geometry_metrics.csv", index=False)
    generate_summary_statistics(results_df, output_dir)
    return results_df
