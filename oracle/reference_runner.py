"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference from /root/reference (this container only).

The reference module does ``import pyvista as pv`` at the top
(/root/reference/src/geometry/tract_geom_proc.py:3) and pyvista is not installed here, so a stub
module is placed in ``sys.modules`` first.  Only ``pv.read`` (tract_geom_proc.py:10) touches it; the
stub's ``read`` returns an object carrying ``.points`` (P,3) and ``.lines`` (legacy
``[n, i0..i(n-1), n, ...]``), so the reference's own loader loop, both filters, all 17 metric
functions and the bundle aggregate run exactly as shipped.

/root/reference does not exist on the GPU box.  Nothing under ``-m gpu``, ``smoke()`` or ``bench.py``
imports this file; it is used by ``tests/golden/make_golden.py`` (fixture generation) and by the
CPU tests that pin ``oracle/streamline_oracle.py`` against the real thing when the reference is present.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("TG_REFERENCE_ROOT", "/root/reference")
_REF_FILE = os.path.join(REFERENCE_ROOT, "src", "geometry", "tract_geom_proc.py")


def available() -> bool:
    return os.path.isfile(_REF_FILE)


class _Mesh:
    def __init__(self, points, lines):
        self.points = points
        self.lines = lines


_state = {"module": None, "stub": None}


def load():
    """Import the reference hot-path module with a stub pyvista; cached."""
    if _state["module"] is not None:
        return _state["module"]
    if not available():
        raise FileNotFoundError(f"reference not present at {_REF_FILE}")
    stub = sys.modules.get("pyvista")
    if stub is None or not hasattr(stub, "__tg_stub__"):
        stub = types.ModuleType("pyvista")
        stub.__tg_stub__ = True
        stub._meshes = {}
        stub.read = lambda path: stub._meshes[str(path)]
        sys.modules["pyvista"] = stub
    spec = importlib.util.spec_from_file_location("_tg_reference_tract_geom_proc", _REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _state["module"], _state["stub"] = mod, stub
    return mod


def csr_to_legacy_lines(offsets):
    """CSR offsets -> legacy VTK ``lines`` array with identity connectivity."""
    offsets = np.asarray(offsets, dtype=np.int64)
    n = np.diff(offsets)
    S = len(n)
    out = np.empty(int(offsets[-1]) + S, dtype=np.int64)
    head = offsets[:-1] + np.arange(S, dtype=np.int64)
    out[head] = n
    mask = np.ones(len(out), dtype=bool)
    mask[head] = False
    out[mask] = np.arange(int(offsets[-1]), dtype=np.int64)
    return out


def reference_compute(points, offsets, max_streamlines=None, lines=None):
    """Call the reference's compute_streamline_metrics on in-memory data.

    ``points`` keeps its dtype (the reference's precision follows it); pass float64
    for the canonical parity contract.
    """
    mod = load()
    stub = _state["stub"]
    if lines is None:
        lines = csr_to_legacy_lines(offsets)
    key = f"<mem:{id(points)}:{id(lines)}>"
    stub._meshes[key] = _Mesh(np.asarray(points), np.asarray(lines))
    try:
        return mod.compute_streamline_metrics(key, max_streamlines=max_streamlines)
    finally:
        stub._meshes.pop(key, None)


def reference_pca_eigs(sl):
    return load().pca_eigs(np.asarray(sl))
