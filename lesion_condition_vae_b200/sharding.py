"""Multi-GPU plumbing of the streamline-metrics path: CSR range sharding and the bundle-partial
exchange.

Polylines are independent (every stencil stays inside one polyline), so the tractogram shards by
contiguous CSR ranges, balanced by POINT count, with no data-path collective.  The only
cross-polyline step of the reference is the 13-column nan-mean per bundle
(/root/reference/src/geometry/tract_geom_proc.py:191-210): every rank reduces its own shard to
``B x 27`` partial moments (13 sums, the kept-row count, 13 non-NaN counts), ONE all-gather moves
them (216 bytes per bundle per rank: latency-bound on NVLink/NVSwitch), and every rank adds the
gathered partials in rank order — deterministic and identical on all ranks, which an all-reduce
would not guarantee.

``torch.distributed`` is plumbing only: NCCL for device tensors, gloo for the CPU tests.
"""
from __future__ import annotations

import numpy as np

N_BUNDLE_COLS = 13
PARTIAL_WIDTH = 2 * N_BUNDLE_COLS + 1          # 13 sums | n_streamlines | 13 non-NaN counts


def shard_ranges(offsets, world_size):
    """Split polylines [0, S) into ``world_size`` contiguous ranges holding ~P/world_size points each.

    Returns int64[world_size + 1] polyline boundaries (non-decreasing, first 0, last S)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    S = len(offsets) - 1
    P = int(offsets[-1]) - int(offsets[0])
    targets = int(offsets[0]) + (np.arange(1, world_size, dtype=np.float64) * P / world_size)
    cuts = np.searchsorted(offsets, targets, side="left").astype(np.int64)
    cuts = np.clip(cuts, 0, S)
    bounds = np.concatenate([[0], cuts, [S]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def shard_csr(points, offsets, lo, hi):
    """The CSR slice of polylines [lo, hi): (points view, rebased offsets copy)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    p0, p1 = int(offsets[lo]), int(offsets[hi])
    return points[p0:p1], offsets[lo:hi + 1] - p0


def shard_bundle_offsets(bundle_offsets, lo, hi):
    """Bundle table of the shard [lo, hi): every bundle keeps its index, clipped to the shard
    (bundles outside it become empty), rebased to the shard's first polyline."""
    bo = np.asarray(bundle_offsets, dtype=np.int64)
    return np.clip(bo, lo, hi) - lo


def pack_partials(sums, counts):
    """(B,13) float64 sums + (B,14) int64 counts -> (B,27) float64 (counts < 2^53 are exact)."""
    sums = np.asarray(sums, dtype=np.float64).reshape(-1, N_BUNDLE_COLS)
    counts = np.asarray(counts, dtype=np.int64).reshape(-1, N_BUNDLE_COLS + 1)
    return np.concatenate([sums, counts.astype(np.float64)], axis=1)


def combine_partials(gathered):
    """(world, B, 27) gathered partials -> (sums (B,13), counts (B,14)), added in rank order."""
    g = np.asarray(gathered, dtype=np.float64)
    total = np.zeros(g.shape[1:], dtype=np.float64)
    for r in range(g.shape[0]):                      # fixed order: same result on every rank
        total = total + g[r]
    sums = total[:, :N_BUNDLE_COLS]
    counts = np.rint(total[:, N_BUNDLE_COLS:]).astype(np.int64)
    return sums, counts


def means_from_partials(sums, counts):
    """n_streamlines (B,) and the 13 nan-means (B,13): sum / non-NaN count, NaN where the count is 0."""
    with np.errstate(invalid="ignore", divide="ignore"):
        means = np.where(counts[:, 1:] > 0, sums / np.maximum(counts[:, 1:], 1), np.nan)
    return counts[:, 0].copy(), means


def allgather_partials(partial, group=None):
    """All-gather a (B,27) float64 torch tensor (CPU with gloo, CUDA with NCCL) -> (world, B, 27)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    flat = partial.contiguous().view(-1)
    out = torch.empty(world * flat.numel(), dtype=partial.dtype, device=partial.device)
    if partial.is_cuda:
        dist.all_gather_into_tensor(out, flat, group=group)
    else:                                            # gloo has no flat all-gather: gather into views
        dist.all_gather(list(out.view(world, -1).unbind(0)), flat, group=group)
    return out.view((world,) + tuple(partial.shape))


class DeviceShard:
    """One rank's share of a device-resident CSR tractogram: polylines [lo, hi) of the global
    table, cut by :func:`shard_ranges` so that every rank holds ~P/world points.

    ``step()`` = tg_metrics_csr_dev on the slice, tg_bundle_partials_dev (kernel 2 writes the ``B x 27`` payload
    itself) and ONE all-gather of that block, all enqueued back to back: no host round trip and no torch kernel
    between kernel 2 and the collective.  ``out`` / ``keep`` hold this rank's slice of the 17 x S table; they stay
    on the device (df_sl is gathered only when a caller wants the DataFrame)."""

    def __init__(self, ctx, points, offsets, lo, hi, bundle_offsets=None, group=None, copy=False):
        import torch
        self.ctx, self.group = ctx, group
        self.lo, self.hi = int(lo), int(hi)
        p0, p1 = int(offsets[self.lo]), int(offsets[self.hi])
        self.S = self.hi - self.lo
        # The slice carries readable slack on both sides — LEAD points in front (the previous shard's last points) and
        # PAD behind (the next shard's first points), zeros where the table ends.  The streaming kernel stages whole
        # 32-byte sectors / 16-byte pieces and sends a polyline whose staging would leave the array to its exact,
        # differently rounded, path; with the slack no shard boundary does that, so every polyline's row is
        # bit-identical to the single-GPU one.  P counts the padded array; offsets are rebased by LEAD.
        LEAD, PAD = (4 if p0 > 0 else 0), 2                # 4 points = 96 bytes: keeps 32-byte alignment classes
        total = int(points.shape[0])
        if not copy and p0 >= LEAD and p1 + PAD <= total:
            self.points = points[p0 - LEAD:p1 + PAD]                        # a view of the global table
        else:
            a, b = max(p0 - LEAD, 0), min(p1 + PAD, total)
            buf = torch.zeros((LEAD + (p1 - p0) + PAD, 3), dtype=points.dtype, device=points.device)
            buf[LEAD - (p0 - a):LEAD + (b - p0)] = points[a:b]
            self.points = buf                                               # an own copy: the global table can be freed
        self.offsets = (offsets[self.lo:self.hi + 1] - (p0 - LEAD)).contiguous()
        self.P = LEAD + (p1 - p0) + PAD
        dev = points.device
        S_all = int(offsets.numel()) - 1
        bo = np.array([0, S_all], dtype=np.int64) if bundle_offsets is None else np.asarray(bundle_offsets, dtype=np.int64)
        self.bundle_offsets = shard_bundle_offsets(bo, self.lo, self.hi)
        self.B = len(bo) - 1
        self.out = torch.empty((17, max(self.S, 1)), dtype=torch.float64, device=dev)
        self.keep = torch.empty(max(self.S, 1), dtype=torch.uint8, device=dev)
        self.partial = torch.zeros((self.B, PARTIAL_WIDTH), dtype=torch.float64, device=dev)

    def compute(self, stream=0):
        """Kernels only (metrics + packed bundle partials) on ``stream``."""
        from . import _lib
        if self.S > 0:
            self.ctx.metrics_dev(self.points.data_ptr(), _lib.F64, self.offsets.data_ptr(), self.S, self.P,
                                 self.out.data_ptr(), self.keep.data_ptr(), stream)
        self.ctx.bundle_partials_dev(self.out.data_ptr(), self.keep.data_ptr(), 0, self.S, self.bundle_offsets,
                                     self.partial.data_ptr(), stream)

    def step(self, stream=0):
        """compute() + the all-gather; returns the (world, B, 27) tensor every rank then sums in rank order."""
        self.compute(stream)
        return allgather_partials(self.partial, self.group)
