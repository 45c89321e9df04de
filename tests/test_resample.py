"""Arc-length resampling to 100 nodes.  The reference has no producer for this step, so
the oracle (oracle/resample_oracle.py) is PARITY-UNPINNED: it restates the published algorithm twice (the
per-node walk and a vectorised form) and the CPU tests check the two against each other and against
analytic cases; the GPU tests compare the CUDA kernel with it at 1e-9 and through size-independent
properties (end points, equal arc-length spacing, nodes lie on the polyline, rigid-motion equivariance)."""
import numpy as np
import pytest

from lesion_condition_vae_b200 import synth
from oracle import resample_oracle as ro


def _cases():
    rng = np.random.default_rng(42)
    n = np.concatenate([synth.lengths_uniform(rng, 60, 2, 160), [2, 3, 33, 64, 65, 700]])
    pts, off = synth.random_walk_csr(n, seed=42)
    lines = [pts[off[i]:off[i + 1]] for i in range(len(n))]
    t = np.linspace(0, 1, 17)
    lines += [
        np.outer(t, [3.0, 4.0, 12.0]),                                            # straight, uniform
        np.outer(t ** 2, [1.0, 0.0, 0.0]),                                        # straight, non-uniform spacing
        np.array([[0, 0, 0], [1, 0, 0], [1, 0, 0], [1, 0, 0], [1, 2, 0.0]]),      # repeated points inside
        np.array([[0, 0, 0], [0, 0, 0], [5, 0, 0.0]]),                            # repeated first point
        np.array([[0, 0, 0], [5, 0, 0], [5, 0, 0.0]]),                            # repeated last point
        np.ones((4, 3)),                                                          # zero length
        np.array([[7.0, 8.0, 9.0]]),                                              # one point
        np.zeros((0, 3)),                                                         # empty
        np.array([[0, 0, 0], [1, np.nan, 0], [2, 0, 0.0]]),                       # non-finite
    ]
    return lines


def test_two_statements_of_the_algorithm_agree():
    for K in (2, 5, 100):
        for line in _cases():
            a, b = ro.resample_walk(line, K), ro.resample_line(line, K)
            assert np.array_equal(np.isnan(a), np.isnan(b))
            scale = max(1.0, float(np.nanmax(np.abs(line))) if line.size and np.isfinite(line).any() else 1.0)
            np.testing.assert_allclose(np.nan_to_num(a), np.nan_to_num(b), rtol=0, atol=1e-13 * scale)


def test_oracle_analytic_cases():
    t = np.linspace(0, 1, 17)
    got = ro.resample_line(np.outer(t ** 2, [2.0, 0.0, 0.0]), 100)                # any parametrisation of a segment
    np.testing.assert_allclose(got, np.outer(np.linspace(0, 1, 100), [2.0, 0, 0]), atol=1e-15)
    got = ro.resample_line(np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0.0]]), 5)      # L-shape, L = 2, step 0.5
    np.testing.assert_allclose(got, [[0, 0, 0], [.5, 0, 0], [1, 0, 0], [1, .5, 0], [1, 1, 0]], atol=1e-15)
    assert np.array_equal(ro.resample_line(np.ones((4, 3)), 7), np.ones((7, 3)))
    assert np.isnan(ro.resample_line(np.zeros((0, 3)), 7)).all()
    line = _cases()[3]
    got = ro.resample_line(line, 100)
    assert np.array_equal(got[0], line[0]) and np.array_equal(got[-1], line[-1])
    np.testing.assert_allclose(np.linalg.norm(np.diff(got, axis=0), axis=1).max(), ro._cumlen(line)[-1] / 99, rtol=1e-3)


def _distance_to_polyline(q, line):
    a, b = line[:-1], line[1:]
    ab = b - a
    den = np.maximum((ab * ab).sum(1), 1e-300)
    u = np.clip(((q[:, None, :] - a[None]) * ab[None]).sum(2) / den[None], 0, 1)
    proj = a[None] + u[..., None] * ab[None]
    return np.sqrt(((q[:, None, :] - proj) ** 2).sum(2)).min(1)


@pytest.mark.gpu
def test_gpu_resample_matches_oracle(gpu_ctx):
    lines = _cases()
    pts, off = synth.lines_to_csr(lines)
    for K in (2, 5, 100, 257):
        got = gpu_ctx.resample_host(pts, off, K)
        exp = ro.resample_csr(pts, off, K)
        assert np.array_equal(np.isnan(got), np.isnan(exp)), K
        for s, line in enumerate(lines):
            if not len(line) or not np.isfinite(line).all():
                continue
            scale = max(1.0, float(np.abs(line).max()))
            np.testing.assert_allclose(got[s], exp[s], rtol=0, atol=1e-9 * scale, err_msg=f"K={K} line {s}")
            assert np.array_equal(got[s, 0], line[0]) and np.array_equal(got[s, -1], line[-1])
    # float32 storage: exact upcast, same nodes as the float64 copy of the same values
    p32 = pts.astype(np.float32)
    a = gpu_ctx.resample_host(p32, off, 100)
    b = gpu_ctx.resample_host(p32.astype(np.float64), off, 100)
    assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.gpu
def test_gpu_resample_argument_checks_and_long_lines(gpu_ctx):
    from lesion_condition_vae_b200 import _lib
    pts, off = synth.random_walk_csr(np.array([5, 130, 129, 1000, 4097, 2, 1]), seed=8)
    with pytest.raises(_lib.TractGeomError):
        gpu_ctx.resample_host(pts, off, 1)                        # fewer than 2 nodes
    with pytest.raises(_lib.TractGeomError):
        gpu_ctx.resample_host(pts, off[::-1].copy(), 10)          # decreasing offsets
    assert gpu_ctx.resample_host(pts[:0], np.zeros(1, np.int64), 10).shape == (0, 10, 3)
    for K in (100, 128, 129, 1000):                                # beyond 128 nodes / 129 points: the generic path
        got = gpu_ctx.resample_host(pts, off, K)
        np.testing.assert_allclose(got, ro.resample_csr(pts, off, K), rtol=0, atol=1e-9 * np.abs(pts).max())
    # an unaligned view of the point array (8 bytes off a 16-byte boundary) and its aligned copy agree bit for bit
    raw = np.zeros(pts.size + 1)
    view = raw[1:].reshape(-1, 3)
    view[:] = pts
    assert np.array_equal(gpu_ctx.resample_host(view, off, 100), gpu_ctx.resample_host(pts, off, 100))


@pytest.mark.gpu
def test_gpu_resample_properties_at_scale(gpu_ctx):
    """200k polylines on the device (sizes the oracle cannot cover): equal arc-length spacing, nodes on the
    polyline, end points exact, rigid-motion equivariance; a subsample against the oracle."""
    import torch
    from lesion_condition_vae_b200 import _lib
    dev = torch.device("cuda:0")
    S, K = 200_000, 100
    n = synth.torch_lengths("normal", S, 9, dev)
    pts, off = synth.torch_random_walk_csr(n, 9, dev)
    P = pts.shape[0]
    nodes = torch.empty((S, K, 3), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.resample_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, K, nodes.data_ptr())
    gpu_ctx.synchronize()
    assert bool(torch.isfinite(nodes).all())
    first, last = pts[off[:-1]], pts[off[1:] - 1]
    assert torch.equal(nodes[:, 0], first) and torch.equal(nodes[:, -1], last)
    # chord between consecutive nodes <= arc step, and close to it for these smooth curves
    seg = torch.linalg.norm(pts[1:] - pts[:-1], dim=1)
    seg[off[1:-1] - 1] = 0.0                                     # joints between polylines
    cum = torch.cat([torch.zeros(1, dtype=torch.float64, device=dev), torch.cumsum(seg, 0)])
    L = cum[off[1:] - 1] - cum[off[:-1]]
    chord = torch.linalg.norm(nodes[:, 1:] - nodes[:, :-1], dim=2)
    step = (L / (K - 1))[:, None]
    assert bool((chord <= step * (1 + 1e-9)).all()) and bool((chord >= step * 0.95).all())
    # translation + rotation equivariance
    th = 0.7
    R = torch.tensor([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1.0]], dtype=torch.float64, device=dev)
    shift = torch.tensor([10.0, -20.0, 5.0], dtype=torch.float64, device=dev)
    moved = torch.empty_like(nodes)
    p2 = (pts @ R.T + shift).contiguous()
    torch.cuda.synchronize()                                     # the context launches on its own stream
    gpu_ctx.resample_dev(p2.data_ptr(), _lib.F64, off.data_ptr(), S, P, K, moved.data_ptr())
    gpu_ctx.synchronize()
    assert float((moved - (nodes @ R.T + shift)).abs().max()) < 1e-9
    idx = np.random.default_rng(3).integers(0, S, 300)
    off_h = off.cpu().numpy()
    for s in idx:
        line = pts[off_h[s]:off_h[s + 1]].cpu().numpy()
        got = nodes[s].cpu().numpy()
        np.testing.assert_allclose(got, ro.resample_line(line, K), rtol=0, atol=1e-9 * max(1.0, np.abs(line).max()))
        assert _distance_to_polyline(got, line).max() < 1e-9
