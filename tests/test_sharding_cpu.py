"""CPU suite, part 3: the N > 1 path on the host side — CSR range sharding balanced by points,
per-shard bundle partials, ONE all-gather (gloo, world_size 2, real processes) and the rank-ordered
combine.  The per-shard arithmetic is done by the oracle here (no GPU in this container); on the
GPU box bench.py feeds the same functions with the kernels' partials over NCCL."""
import os
import socket

import numpy as np
import pytest

from lesion_condition_vae_b200 import sharding, synth
from oracle import streamline_oracle as so

SRC = (0, 2, 4, 6, 7, 8, 10, 11, 12, 16, 13, 14, 15)      # df_sl column feeding bundle column j


def _oracle_partials(points, offsets, bundle_offsets):
    """(B,13) sums and (B,14) counts of a shard, from the oracle's per-streamline table."""
    B = len(bundle_offsets) - 1
    sums = np.zeros((B, 13)); counts = np.zeros((B, 14), np.int64)
    table, src = so.per_streamline_table(points, offsets)
    for b in range(B):
        rows = table[(src >= bundle_offsets[b]) & (src < bundle_offsets[b + 1])]
        counts[b, 0] = len(rows)
        if len(rows):
            cols = rows[:, SRC]
            ok = ~np.isnan(cols)
            sums[b] = np.where(ok, cols, 0.0).sum(axis=0)
            counts[b, 1:] = ok.sum(axis=0)
    return sums, counts


def test_shard_ranges_balance_points():
    rng = np.random.default_rng(0)
    n = synth.lengths_heavy_tail(rng, 5000, 10, 5000)
    off = synth.offsets_from_lengths(n)
    for world in (1, 2, 3, 8):
        b = sharding.shard_ranges(off, world)
        assert b[0] == 0 and b[-1] == len(n) and np.all(np.diff(b) >= 0) and len(b) == world + 1
        pts = np.diff(off[b])
        assert pts.sum() == off[-1]
        assert pts.max() - pts.min() <= 2 * 5000          # within one (longest) polyline of perfect balance
    assert sharding.shard_ranges(np.zeros(1, np.int64), 4).tolist() == [0, 0, 0, 0, 0]


def test_shard_bundle_offsets_clip():
    bo = np.array([0, 10, 10, 25, 40])
    assert sharding.shard_bundle_offsets(bo, 12, 30).tolist() == [0, 0, 0, 13, 18]


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts, off, bo = synth.config2(S=40, n_tracts=3, n_tp=2)          # 6 bundles x 40 polylines
        bounds = sharding.shard_ranges(off, world)
        lo, hi = int(bounds[rank]), int(bounds[rank + 1])
        p, o = sharding.shard_csr(pts, off, lo, hi)
        sb = sharding.shard_bundle_offsets(bo, lo, hi)
        sums, counts = _oracle_partials(p, o, sb)
        part = torch.from_numpy(sharding.pack_partials(sums, counts))
        gathered = sharding.allgather_partials(part).numpy()
        tsums, tcounts = sharding.combine_partials(gathered)
        n_sl, means = sharding.means_from_partials(tsums, tcounts)
        np.savez(os.path.join(tmp, f"r{rank}.npz"), n=n_sl, means=means, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_allgather_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert r0["lo"] == 0 and r0["hi"] == r1["lo"] and r1["hi"] == 240
    assert np.array_equal(r0["n"], r1["n"]) and np.array_equal(r0["means"], r1["means"], equal_nan=True)   # identical on all ranks
    pts, off, bo = synth.config2(S=40, n_tracts=3, n_tp=2)
    for b in range(len(bo) - 1):
        p, o = sharding.shard_csr(pts, off, int(bo[b]), int(bo[b + 1]))
        _, ref_b = so.compute_streamline_metrics_csr(p, o)
        ref = ref_b.iloc[0].to_numpy(float)
        assert r0["n"][b] == ref[0]                                   # counts bit-exact
        np.testing.assert_allclose(r0["means"][b], ref[1:], rtol=1e-12, atol=0)


def test_device_shard_slices_carry_readable_slack():
    """sharding.DeviceShard (host-side logic, CPU tensors, no context): every slice keeps its polylines' points at
    points[offsets[s]], carries 4 points of lead (when it is not the first) and 2 of pad, as a view of the global table
    when the neighbours exist and as a zero-padded copy otherwise; bundle tables are clipped and rebased."""
    import torch
    from lesion_condition_vae_b200 import sharding
    rng = np.random.default_rng(3)
    n = rng.integers(0, 9, size=200)
    off = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    pts = rng.normal(size=(int(off[-1]), 3))
    t_pts, t_off = torch.from_numpy(pts), torch.from_numpy(off)
    bo = np.array([0, 50, 50, 170, 200], dtype=np.int64)
    bounds = sharding.shard_ranges(off, 4)
    covered = 0
    for r in range(4):
        for copy in (False, True):
            sh = sharding.DeviceShard(None, t_pts, t_off, bounds[r], bounds[r + 1], bo, copy=copy)
            lo, hi = int(bounds[r]), int(bounds[r + 1])
            assert sh.S == hi - lo and sh.points.shape[0] == sh.P
            o = sh.offsets.numpy()
            lead = int(o[0])
            assert lead == (4 if off[lo] > 0 else 0) and sh.P == lead + int(off[hi] - off[lo]) + 2
            for s in (0, sh.S // 2, sh.S - 1):
                if sh.S:
                    a, b = int(o[s]), int(o[s + 1])
                    assert np.array_equal(sh.points[a:b].numpy(), pts[off[lo + s]:off[lo + s + 1]])
            is_view = sh.points.untyped_storage().data_ptr() == t_pts.untyped_storage().data_ptr()
            assert is_view == (not copy and off[lo] >= lead and off[hi] + 2 <= len(pts))
            assert np.array_equal(sh.bundle_offsets, np.clip(bo, lo, hi) - lo)
        covered += hi - lo
    assert covered == 200
