#!/usr/bin/env python3
"""Golden CSVs of the batched driver: the UNMODIFIED reference driver
(/root/reference/src/geometry/comprehensive_tract_geometry_analysis.py::process_all_tracts) run on
the synthetic study tree of tests/dataset_fixture.py, in this container.

    python tests/golden/make_driver_golden.py     # rewrites tests/golden/driver_ms100.csv (the shipped max_streamlines=100) and driver_ms5.csv

The reference imports `tract_geom_proc` flat and `pyvista` at module top; /root/reference/src/geometry
goes on sys.path and a stub pyvista whose read() parses the file with this repo's VTK reader is
installed (the reference decompresses .gz itself, so the stub only ever sees .vtk).  Points keep the
dtype stored in the file, as PyVista would return them; float files are upcast to float64 HERE so
that the golden numbers are the float64 contract."""
import contextlib
import io
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from lesion_condition_vae_b200 import vtk_io  # noqa: E402
from oracle.reference_runner import csr_to_legacy_lines  # noqa: E402
import dataset_fixture  # noqa: E402


class _Mesh:
    def __init__(self, path):
        pts, off = vtk_io.read_polylines_csr(path)
        self.points = np.asarray(pts, dtype=np.float64)
        self.lines = csr_to_legacy_lines(off)


def main():
    stub = types.ModuleType("pyvista")
    stub.read = lambda p: _Mesh(str(p))
    sys.modules["pyvista"] = stub
    sys.path.insert(0, "/root/reference/src/geometry")
    import comprehensive_tract_geometry_analysis as ref
    with tempfile.TemporaryDirectory() as tmp:
        data, cfg = dataset_fixture.build(tmp)
        import json
        config = json.load(open(cfg))
        for ms, name in ((100, "driver_ms100.csv"), (5, "driver_ms5.csv")):
            with contextlib.redirect_stdout(io.StringIO()):
                df = ref.process_all_tracts(config, Path(data), Path(tmp) / "out", max_streamlines=ms)
            df.to_csv(os.path.join(HERE, name), index=False)
            print(name, df.shape, list(df.columns)[:3], "...", list(df.columns)[-4:])


if __name__ == "__main__":
    main()
