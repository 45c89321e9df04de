"""TEST INFRASTRUCTURE — CPU oracle for the arc-length resampling step.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this file; the product path
(lesion_condition_vae_b200/, libtractgeom.so) never does.

PARITY UNPINNED.  /root/reference ships NO producer for the 100-node tables that
src/vae/data_loader.py:94-100 consumes (it only checks `len(nodes) == 100`), so there is no reference
code, test or fixture to pin this oracle to.  It restates the published algorithm tractography tools
use for that step (dipy.tracking.streamline.set_number_of_points, dipy >= 1.0; dipy is not a dependency
of the reference and is not installed here):

    arclengths[0] = 0, arclengths[i] = arclengths[i-1] + |p_i - p_{i-1}|
    step = arclengths[n-1] / (K - 1); node k at arc length k * step, linearly interpolated inside the
    segment that contains it; the last node is the last point itself.

Two statements of it are kept and checked against each other: a per-node Python walk (the published loop,
small cases) and a vectorised numpy one (np.searchsorted).  Conventions where the published code is silent:
a zero-length polyline repeats its first point, an empty one or one with a non-finite length gives NaN.
"""
import numpy as np


def _cumlen(line):
    d = np.diff(line, axis=0)
    seg = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
    return np.concatenate([[0.0], np.cumsum(seg)])


def resample_walk(line, K=100):
    """The published loop: walk the segments once, emitting nodes as their arc length is passed."""
    line = np.asarray(line, dtype=np.float64).reshape(-1, 3)
    n = len(line)
    out = np.full((K, 3), np.nan)
    if n == 0:
        return out
    cum = _cumlen(line)
    L = cum[-1]
    if not np.isfinite(L):
        return out
    if L == 0.0:
        out[:] = line[0]
        return out
    step = L / (K - 1)
    j = 1
    for k in range(K - 1):
        t = k * step
        while j < n - 1 and not (t < cum[j]):
            j += 1
        c0, c1 = cum[j - 1], cum[j]
        r = min(max((t - c0) / (c1 - c0), 0.0), 1.0)
        out[k] = line[j - 1] + r * (line[j] - line[j - 1])
    out[K - 1] = line[n - 1]
    return out


def resample_line(line, K=100):
    """Vectorised statement of the same algorithm."""
    line = np.asarray(line, dtype=np.float64).reshape(-1, 3)
    n = len(line)
    out = np.full((K, 3), np.nan)
    if n == 0:
        return out
    cum = _cumlen(line)
    L = cum[-1]
    if not np.isfinite(L):
        return out
    if L == 0.0:
        out[:] = line[0]
        return out
    t = np.arange(K - 1) * (L / (K - 1))
    j = np.clip(np.searchsorted(cum, t, side="right") - 1, 0, n - 2)      # last segment start with cum <= t
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.clip((t - cum[j]) / (cum[j + 1] - cum[j]), 0.0, 1.0)
    out[:K - 1] = line[j] + r[:, None] * (line[j + 1] - line[j])
    out[K - 1] = line[n - 1]
    return out


def resample_csr(points, offsets, K=100):
    points = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    offsets = np.asarray(offsets, dtype=np.int64)
    S = len(offsets) - 1
    out = np.empty((S, K, 3))
    for s in range(S):
        out[s] = resample_line(points[offsets[s]:offsets[s + 1]], K)
    return out
