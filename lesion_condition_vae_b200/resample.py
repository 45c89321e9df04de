"""Along-tract resampling to a fixed number of nodes.

The reference's VAE loader (/root/reference/src/vae/data_loader.py:94-100) reads per-point tables with
exactly 100 ``point_id``s per ``streamline_id`` and a ``position_along_tract`` column (:127-128), but the
repository holds no code that produces them from the ragged tract files.  This module is that producer's
geometric half: every polyline of a CSR tractogram is resampled at 100 (or ``n_nodes``) equally spaced arc
lengths by ``k_resample`` in libtractgeom.so (one warp per polyline).  Sampling image volumes at the nodes
(FA, MD, lesion masks) is outside this path.

No CPU fallback: the call raises if the library or a device is missing.
"""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import _lib, vtk_io

N_NODES = 100          # data_loader.py:97 accepts nothing else


def resample_streamlines_csr(points, offsets, n_nodes: int = N_NODES, ctx=None):
    """-> (nodes float64 (S, n_nodes, 3), position_along_tract float64 (n_nodes,) = linspace(0, 1))."""
    ctx = ctx or _lib.default_context()
    nodes = ctx.resample_host(points, offsets, n_nodes)
    return nodes, np.linspace(0.0, 1.0, n_nodes)


def resample_vtk(vtk_path: str, n_nodes: int = N_NODES, ctx=None):
    """Resample the polylines of a legacy VTK tract file (same reader as compute_streamline_metrics)."""
    points, offsets = vtk_io.read_polylines_csr(vtk_path)
    return resample_streamlines_csr(points, offsets, n_nodes, ctx)


def nodes_to_long_frame(nodes, position=None, tract_id=None):
    """Long table in the layout data_loader.py:63-96 reads: one row per (streamline_id, point_id), with
    ``position_along_tract`` and the node coordinates (the loader's feature columns are added by whoever
    samples the image volumes at x, y, z)."""
    nodes = np.asarray(nodes)
    S, K, _ = nodes.shape
    if position is None:
        position = np.linspace(0.0, 1.0, K)
    df = pd.DataFrame({
        "streamline_id": np.repeat(np.arange(S, dtype=np.int64), K),
        "point_id": np.tile(np.arange(K, dtype=np.int64), S),
        "position_along_tract": np.tile(np.asarray(position, dtype=np.float64), S),
        "x": nodes[:, :, 0].reshape(-1), "y": nodes[:, :, 1].reshape(-1), "z": nodes[:, :, 2].reshape(-1),
    })
    if tract_id is not None:
        df.insert(0, "tract_id", tract_id)
    return df
