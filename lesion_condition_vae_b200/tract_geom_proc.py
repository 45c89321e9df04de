"""Host-side mirror of the reference module ``tract_geom_proc`` for the one function this repo
replaces:

    compute_streamline_metrics(vtk_path, max_streamlines=None) -> (df_sl, df_bundle)
    /root/reference/src/geometry/tract_geom_proc.py:153-212

Same name, same signature, same DataFrames (17 float64 columns / 1 x 14 summary), same errors
(missing file -> FileNotFoundError from the reader; nothing survives -> KeyError('length'),
SURVEY.md N4).  The arithmetic runs in libtractgeom.so on a B200; there is no CPU fallback — if
the library or a device is missing the call raises.

Points are upcast exactly to float64 on the device when the file stores float32; the reference
would compute those files mostly in float32, so on float32 files this path is
the more accurate one and differs from the literal reference by ~1e-7.  On float64 points every
metric matches the reference to 1e-9 relative (tests/).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import pandas as pd

from . import _lib, vtk_io

SL_COLUMNS = (  # ref:164-187
    "length", "end_to_end", "tortuosity", "straightness", "curv_mean", "curv_std", "curv_energy",
    "torsion_mean", "bend_angle_mean", "bbox_vol", "elongation_ratio", "planarity_ratio",
    "anisotropy_ratio", "centroid_x", "centroid_y", "centroid_z", "ang_dispersion",
)
BUNDLE_COLUMNS = (  # ref:196-209
    "n_streamlines", "length_mean", "tortuosity_mean", "curv_mean_avg", "curv_energy_mean",
    "torsion_mean_avg", "bend_angle_mean_avg", "elongation_ratio_mean", "planarity_ratio_mean",
    "anisotropy_ratio_mean", "ang_dispersion_mean", "centroid_x_mean", "centroid_y_mean", "centroid_z_mean",
)


def read_streamlines_from_vtk(vtk_path: str, max_streamlines: Optional[int] = None) -> List[np.ndarray]:
    """ref:9-26 — list of (n,3) arrays that pass the loader filter, at most ``max_streamlines``.

    Kept for API compatibility; :func:`compute_streamline_metrics` does not go through a Python
    list (the filter is evaluated on the device and reported in the ``keep`` flags)."""
    pts, off = vtk_io.read_polylines_csr(vtk_path)
    out = []
    for s in range(len(off) - 1):
        sl = pts[off[s]:off[s + 1]]
        if sl.shape[0] > 2 and np.isfinite(sl).all():
            out.append(sl)
            if max_streamlines is not None and len(out) >= max_streamlines:
                break
    return out


BUNDLE_SOURCE = (  # df_sl column behind each of the 13 means
    "length", "tortuosity", "curv_mean", "curv_energy", "torsion_mean", "bend_angle_mean", "elongation_ratio",
    "planarity_ratio", "anisotropy_ratio", "ang_dispersion", "centroid_x", "centroid_y", "centroid_z",
)
# opt-in extra df_bundle columns, appended AFTER the reference's 14 so the default schema
# read by classification.py:64-75 / correlation.py:70-75 never changes
SPREAD_COLUMNS = tuple(f"{src}_{stat}" for src in BUNDLE_SOURCE for stat in ("std", "min", "max"))


def _prefix_for(n_per_line, want, start=0):
    """Smallest end index e > start such that lines[start:e] holds `want` lines with n > 2."""
    cand = np.flatnonzero(n_per_line[start:] > 2)
    if len(cand) < want:
        return len(n_per_line)
    return start + int(cand[want - 1]) + 1


def streamline_table_csr(points, offsets, max_streamlines=None, ctx=None, spread=None):
    """-> (out (17,S') float64, row_mask bool[S'], sums (13,), counts (14,)) for the processed prefix.

    ``row_mask`` marks the polylines that become df_sl rows.  With ``max_streamlines`` the
    reference stops reading after that many polylines passed the LOADER filter and applies the
    length filter afterwards; the prefix that must be processed is found from the
    point counts and extended only if some candidate turns out to hold a non-finite coordinate.
    """
    ctx = ctx or _lib.default_context()
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    points = np.asarray(points)
    S = len(offsets) - 1
    if max_streamlines is None:
        out, keep, sums, counts = ctx.metrics_host(points, offsets, spread=spread)
        return out, keep == _lib.KEEP_BOTH, sums[0], counts[0]
    want = int(max_streamlines)
    # ref:22-24: the cap is tested after an append, so a cap <= 0 still admits one polyline
    want = max(want, 1)
    n_per_line = np.diff(offsets)
    end = _prefix_for(n_per_line, want)
    while True:
        # one bundle = the whole prefix, so the device reduction already covers exactly the rows
        out, keep, sums, counts = ctx.metrics_host(points[:int(offsets[end])], offsets[:end + 1], spread=spread)
        have = int(((keep & _lib.KEEP_LOADER) != 0).sum())
        if have >= want or end >= S:
            break
        end = _prefix_for(n_per_line, want - have, start=end)     # some candidates were non-finite
    # the prefix ends ON the want-th candidate, so have <= want and every loader-accepted polyline
    # of the prefix is one the reference would have read
    return out, keep == _lib.KEEP_BOTH, sums[0], counts[0]


def frames_from_table(out, rows, sums, counts, spread=None):
    """Build (df_sl, df_bundle) exactly as ref:189-211 shapes them (+ the opt-in spread columns when
    ``spread`` (13,3) is given)."""
    n_rows = int(counts[0])
    if n_rows == 0:
        # ref:189,197: pd.DataFrame([]) has no columns, df_sl["length"] raises KeyError('length')
        raise KeyError("length")
    # one (rows, 17) float64 block that the frame owns (3x cheaper to build than 17 column arrays: at 1,000 polylines the
    # frames used to cost more than the device call)
    table = out.T.copy() if rows.all() else np.ascontiguousarray(out[:, rows].T)
    df_sl = pd.DataFrame(table, columns=list(SL_COLUMNS), copy=False)
    with np.errstate(invalid="ignore", divide="ignore"):
        means = np.where(counts[1:] > 0, sums / np.maximum(counts[1:], 1), np.nan)
    bundle = {"n_streamlines": np.array([n_rows], dtype=np.int64)}
    for name, v in zip(BUNDLE_COLUMNS[1:], means):
        bundle[name] = np.array([v], dtype=np.float64)
    if spread is not None:
        for name, v in zip(SPREAD_COLUMNS, np.asarray(spread, dtype=np.float64).reshape(-1)):
            bundle[name] = np.array([v], dtype=np.float64)
    return df_sl, pd.DataFrame(bundle)


def compute_streamline_metrics_csr(points, offsets, max_streamlines: Optional[int] = None, ctx=None, extra_stats=False):
    """Same contract as :func:`compute_streamline_metrics`, on an in-memory CSR tractogram.

    ``extra_stats=True`` appends 39 opt-in columns to df_bundle: np.nanstd / np.nanmin / np.nanmax of the 13
    aggregated df_sl columns (``SPREAD_COLUMNS``), reduced on the device in the same call."""
    spread = np.empty((1, _lib.N_BUNDLE_COLS, 3), dtype=np.float64) if extra_stats else None
    out, rows, sums, counts = streamline_table_csr(points, offsets, max_streamlines, ctx, spread=spread)
    return frames_from_table(out, rows, sums, counts, None if spread is None else spread[0])


def _load_for_device(vtk_path):
    """ref:9-26 without the Python loop: the file's POINTS block lands in pinned memory in the file's own storage type
    (big-endian float/double for binary files: the device swaps and upcasts), the cell array becomes CSR offsets."""
    try:
        arena = _lib.default_arena()
        arena.reset()                                      # the previous call's views are dead: results were copied out
    except _lib.TractGeomError:
        arena = None                                       # no device: the call below raises anyway
    return vtk_io.read_polylines_raw(vtk_path, arena)


def compute_streamline_metrics(vtk_path: str, max_streamlines: Optional[int] = None) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """Returns: df_sl (per streamline) and df_bundle (bundle-level summary).  Drop-in for ref:153."""
    points, offsets = _load_for_device(vtk_path)
    return compute_streamline_metrics_csr(points, offsets, max_streamlines)


def compute_streamline_metrics_extended(vtk_path: str, max_streamlines: Optional[int] = None) -> Tuple[pd.DataFrame, pd.DataFrame]:
    """:func:`compute_streamline_metrics` with the opt-in spread columns (``SPREAD_COLUMNS``) appended to
    df_bundle.  A separate name, so the drop-in keeps the reference's exact signature and schema."""
    points, offsets = _load_for_device(vtk_path)
    return compute_streamline_metrics_csr(points, offsets, max_streamlines, extra_stats=True)


def compute_bundles_csr(points, offsets, bundle_offsets, ctx=None, want_rows=True, extra_stats=False):
    """Many bundles in ONE launch (BASELINE config 2: 16 tracts x 4 timepoints).

    Returns a list of (df_sl | None, df_bundle | None) per bundle; a bundle with no surviving
    polyline yields (None, None) — the per-file call would have raised KeyError('length') and the
    reference driver would have skipped it (comprehensive_tract_geometry_analysis.py:129-131)."""
    ctx = ctx or _lib.default_context()
    bo = np.asarray(bundle_offsets, dtype=np.int64)
    spread = np.empty((len(bo) - 1, _lib.N_BUNDLE_COLS, 3), dtype=np.float64) if extra_stats else None
    out, keep, sums, counts = ctx.metrics_host(points, offsets, bo, want_rows=want_rows, spread=spread)
    res = []
    for b in range(len(bo) - 1):
        if counts[b, 0] == 0:
            res.append((None, None))
            continue
        lo, hi = int(bo[b]), int(bo[b + 1])
        if want_rows:
            df_sl, df_b = frames_from_table(out[:, lo:hi], keep[lo:hi] == _lib.KEEP_BOTH, sums[b], counts[b],
                                            None if spread is None else spread[b])
        else:
            _, df_b = frames_from_table(np.empty((17, 0)), np.zeros(0, bool), sums[b], counts[b],
                                        None if spread is None else spread[b])
            df_sl = None
        res.append((df_sl, df_b))
    return res
