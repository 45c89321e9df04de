"""Deterministic synthetic tractograms in CSR form.

A tractogram is ``(points, offsets)``: ``points`` is ``(P, 3)`` float64, ``offsets`` is
``int64[S+1]`` and streamline ``s`` owns rows ``offsets[s]:offsets[s+1]``.  This is the layout
the kernels consume; the reference builds the same thing one streamline at a time from the
legacy VTK ``lines`` array (/root/reference/src/geometry/tract_geom_proc.py:17-25).

Curve model (all configs): a correlated random walk.  ``p0 ~ U(-50,50)^3`` mm, ``u0`` a random
unit vector, ``u[i+1] = normalize(u[i] + sigma*N(0,I))``, ``p[i+1] = p[i] + u[i+1]*step*(1+jitter*U(-1,1))``.
``sigma >= 0.02`` keeps the covariance condition number small enough that the reference's own
LAPACK eigenvalues are good to ~1e-11, so the 1e-9 parity contract is meaningful.

Two back ends produce the *same model* but not the same random stream: numpy (host, used for the
golden fixtures and CPU tests) and torch (any device, used by bench.py for >=1M streamlines).
"""
from __future__ import annotations

import os

import numpy as np

SIGMA = 0.05
STEP_MM = 0.5
JITTER = 0.1

# BASELINE.json configs -> length law
def lengths_uniform(rng, S, lo, hi):
    """n ~ U{lo..hi} inclusive (configs 1 and 2)."""
    return rng.integers(lo, hi + 1, size=S, dtype=np.int64)


def lengths_normal(rng, S, mean=100.0, sd=15.0, lo=3, hi=200):
    """n = clip(round(N(mean, sd^2)), lo, hi) (configs 3 and 5)."""
    return np.clip(np.rint(rng.normal(mean, sd, size=S)), lo, hi).astype(np.int64)


def lengths_heavy_tail(rng, S, nmin=10, nmax=5000):
    """n = min(nmax, floor(nmin/U)), U ~ Uniform(0,1]: truncated Pareto, alpha=1 (config 4)."""
    u = 1.0 - rng.random(S)  # (0, 1]
    return np.minimum(nmax, np.floor(nmin / u)).astype(np.int64)


def lengths_log_uniform(rng, S, nmin=10, nmax=5000):
    """n = floor(nmin (nmax/nmin)^U), U ~ Uniform[0,1): log-uniform on [nmin, nmax], mean ~ 803 (config 4 variant)."""
    return np.minimum(nmax, np.floor(nmin * (nmax / nmin) ** rng.random(S))).astype(np.int64)


def offsets_from_lengths(n):
    off = np.zeros(len(n) + 1, dtype=np.int64)
    np.cumsum(n, out=off[1:])
    return off


def random_walk_csr(lengths, seed, sigma=SIGMA, step=STEP_MM, jitter=JITTER):
    """Host generator.  Vectorised across streamlines, sequential along each one.

    Streamlines are visited longest first so the active set at point index ``i`` is a prefix.
    """
    lengths = np.asarray(lengths, dtype=np.int64)
    S = len(lengths)
    off = offsets_from_lengths(lengths)
    pts = np.empty((int(off[-1]), 3), dtype=np.float64)
    if S == 0:
        return pts, off
    rng = np.random.default_rng(seed)
    order = np.argsort(-lengths, kind="stable")
    n_sorted = lengths[order]
    base = off[:-1][order]
    p = rng.uniform(-50.0, 50.0, size=(S, 3))
    u = rng.normal(size=(S, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    nmax = int(n_sorted[0]) if S else 0
    # active(i) = number of streamlines with n > i
    neg = -n_sorted
    for i in range(nmax):
        act = int(np.searchsorted(neg, -i, side="left"))  # n_sorted > i  <=>  -n_sorted < -i
        if act == 0:
            break
        if i > 0:
            ua = u[:act] + sigma * rng.normal(size=(act, 3))
            ua /= np.linalg.norm(ua, axis=1, keepdims=True)
            u[:act] = ua
            p[:act] += ua * (step * (1.0 + jitter * rng.uniform(-1.0, 1.0, size=(act, 1))))
        pts[base[:act] + i] = p[:act]
    return pts, off


# ---------------------------------------------------------------------------------------------
# Named configurations (BASELINE.json "configs", sized by SURVEY.md §8d)
# ---------------------------------------------------------------------------------------------
TRACT_NAMES = (
    "chip_right", "hipcom", "thalsub_left", "cing_left", "thalsub_right", "cing_right",
    "fimbria_left", "ant_comm", "fimbria_right", "atr_left", "fornix_left", "intcap_left",
    "atr_right", "chip_left", "fornix_right", "intcap_right",
)  # the 16 bundle names of comprehensive_tract_geometry_analysis.py:25-32
TIMEPOINTS = ("2d", "9d", "1mo", "5mo")  # comprehensive_tract_geometry_analysis.py:162


def config1(S=1000, seed=0):
    """One tract, S streamlines, n ~ U{20..119}."""
    rng = np.random.default_rng(seed)
    n = lengths_uniform(rng, S, 20, 119)
    return random_walk_csr(n, seed + 7919)


def config2_bundle(tract_idx, tp_idx, S=5000):
    """One of the 64 bundles of config 2: n ~ U{40..139}, seed 1000 + 4*tract + tp."""
    seed = 1000 + 4 * tract_idx + tp_idx
    rng = np.random.default_rng(seed)
    n = lengths_uniform(rng, S, 40, 139)
    return random_walk_csr(n, seed + 7919)


def config2(S=5000, n_tracts=16, n_tp=4):
    """All bundles of config 2 concatenated: (points, offsets, bundle_offsets)."""
    P, O, B = [], [np.zeros(1, np.int64)], [0]
    base = 0
    for t in range(n_tracts):
        for k in range(n_tp):
            p, o = config2_bundle(t, k, S)
            P.append(p)
            O.append(o[1:] + base)
            base += int(o[-1])
            B.append(B[-1] + len(o) - 1)
    return np.concatenate(P), np.concatenate(O), np.asarray(B, dtype=np.int64)


def adversarial_lines():
    """Small fixed set of edge-case polylines: returns a list of (n,3) arrays.

    Order matters for the max_streamlines tests: index 4 (n=2) and 5 (NaN) are dropped by the
    loader filter, index 6 (zero length) by the L<=1e-8 filter.
    """
    t = np.arange(20, dtype=np.float64)[:, None]
    rng = np.random.default_rng(12345)
    straight = t * np.array([[3.0, 4.0, 12.0]]) / 19.0
    th = np.linspace(0.0, 1.5 * np.pi, 40)
    planar = np.stack([10 * np.cos(th), 10 * np.sin(th), np.zeros_like(th)], axis=1)
    tilted = planar @ np.linalg.qr(rng.normal(size=(3, 3)))[0] + np.array([100.0, -40.0, 7.0])
    three = np.array([[0.0, 0.0, 0.0], [1.0, 0.2, 0.0], [2.0, 0.1, 0.3]])
    two = np.array([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]])
    nanpt = rng.normal(size=(10, 3)); nanpt[4, 1] = np.nan
    zero = np.tile(np.array([[1.5, -2.5, 3.5]]), (5, 1))
    dup = np.cumsum(rng.normal(size=(12, 3)), axis=0); dup[5] = dup[4]; dup[9] = dup[7]
    helix = np.stack([5 * np.cos(th * 2), 5 * np.sin(th * 2), 0.7 * th], axis=1)
    four = np.cumsum(rng.normal(size=(4, 3)), axis=0)
    infpt = rng.normal(size=(6, 3)); infpt[0, 0] = np.inf
    far = np.cumsum(rng.normal(size=(50, 3)) * 0.4, axis=0) + 1000.0
    zigzag = np.stack([np.arange(30.0), (np.arange(30) % 2) * 1.0, np.zeros(30)], axis=1)
    backtrack = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 0], [1, 0, 0], [2, 0.5, 0.1]], dtype=np.float64)
    return [straight, planar, tilted, three, two, nanpt, zero, dup, helix, four, infpt, far, zigzag, backtrack]


def lines_to_csr(lines):
    n = np.array([len(l) for l in lines], dtype=np.int64)
    off = offsets_from_lengths(n)
    pts = np.concatenate([np.asarray(l, dtype=np.float64).reshape(-1, 3) for l in lines]) if len(lines) else np.empty((0, 3))
    return np.ascontiguousarray(pts), off


# ---------------------------------------------------------------------------------------------
# torch back end (device generation for configs 3-5; bench.py)
# ---------------------------------------------------------------------------------------------
def torch_lengths(kind, S, seed, device):
    import torch
    g = torch.Generator(device=device); g.manual_seed(seed)
    if kind == "normal":
        n = torch.randn(S, generator=g, device=device, dtype=torch.float64) * 15.0 + 100.0
        return n.round_().clamp_(3, 200).to(torch.int64)
    if kind == "heavy":
        u = 1.0 - torch.rand(S, generator=g, device=device, dtype=torch.float64)
        return torch.clamp(torch.floor(10.0 / u), max=float(os.environ.get("TG_HEAVY_NMAX", 5000))).to(torch.int64)   # (the env var is a probe: configs[3] is 5000)
    if kind == "loguniform":
        u = torch.rand(S, generator=g, device=device, dtype=torch.float64)
        return torch.clamp(torch.floor(10.0 * torch.pow(torch.tensor(500.0, dtype=torch.float64, device=device), u)), max=5000.0).to(torch.int64)
    if kind == "uniform":                      # BASELINE configs[0]: n ~ U{20..119}
        return torch.randint(20, 120, (S,), generator=g, device=device, dtype=torch.int64)
    if kind == "uniform40":                    # BASELINE configs[1]: n ~ U{40..139}
        return torch.randint(40, 140, (S,), generator=g, device=device, dtype=torch.int64)
    if kind == "fixed96":                      # every polyline 96 points = 36 whole 64-byte atoms (traffic probe)
        return torch.full((S,), 96, device=device, dtype=torch.int64)
    raise ValueError(kind)


def torch_random_walk_csr(lengths, seed, device, sigma=SIGMA, step=STEP_MM, jitter=JITTER):
    """Same model as :func:`random_walk_csr`, on ``device``; returns (points (P,3) f64, offsets i64)."""
    import torch
    S = lengths.numel()
    off = torch.zeros(S + 1, dtype=torch.int64, device=device)
    torch.cumsum(lengths, 0, out=off[1:])
    P = int(off[-1].item())
    pts = torch.empty((P, 3), dtype=torch.float64, device=device)
    if S == 0:
        return pts, off
    g = torch.Generator(device=device); g.manual_seed(seed)
    n_sorted, order = torch.sort(lengths, descending=True, stable=True)
    base = off[:-1][order]
    p = torch.rand((S, 3), generator=g, device=device, dtype=torch.float64) * 100.0 - 50.0
    u = torch.randn((S, 3), generator=g, device=device, dtype=torch.float64)
    u /= torch.linalg.norm(u, dim=1, keepdim=True)
    nmax = int(n_sorted[0].item())
    # active counts for every i in one shot (host side, tiny)
    hist = torch.bincount(n_sorted, minlength=nmax + 1)
    act_ge = torch.flip(torch.cumsum(torch.flip(hist, [0]), 0), [0])  # act_ge[k] = #{n >= k}
    act_host = act_ge.cpu().tolist()
    for i in range(nmax):
        act = act_host[i + 1] if i + 1 <= nmax else 0  # n > i
        if act == 0:
            break
        if i > 0:
            ua = u[:act] + sigma * torch.randn((act, 3), generator=g, device=device, dtype=torch.float64)
            ua /= torch.linalg.norm(ua, dim=1, keepdim=True)
            u[:act] = ua
            r = torch.rand((act, 1), generator=g, device=device, dtype=torch.float64) * 2.0 - 1.0
            p[:act] += ua * (step * (1.0 + jitter * r))
        pts.index_copy_(0, base[:act] + i, p[:act])
    return pts, off
