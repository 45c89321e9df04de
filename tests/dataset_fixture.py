"""Deterministic synthetic study tree for the driver tests (and for tests/golden/make_driver_golden.py):

    <root>/data/<subject>/<timepoint>/bundles/<tract>_curves.vtk[.gz]

4 subjects in 3 groups x 4 timepoints x 16 tracts, a few files missing, a mix of .vtk / .vtk.gz,
binary / ASCII, float / double points, classic / offsets cell layout, and files with polylines
the two filters drop (n <= 2, NaN, zero length) including one file where nothing survives."""
import json
import os

import numpy as np

from lesion_condition_vae_b200 import synth, vtk_io
from lesion_condition_vae_b200.tract_driver import TIMEPOINTS, TRACT_LIST

CONFIG = {"groups": {"Sham": [1017, 1035], "TBI": [1043], "PTE": [1008], "Other": [9999]}, "timepoints": TIMEPOINTS}


def build(root):
    root = str(root)
    data = os.path.join(root, "data")
    k = 0
    for group, subjects in CONFIG["groups"].items():
        for subj in subjects:
            for ti, tp in enumerate(TIMEPOINTS):
                d = os.path.join(data, str(subj), tp, "bundles")
                os.makedirs(d, exist_ok=True)
                for xi, tract in enumerate(TRACT_LIST):
                    k += 1
                    if k % 17 == 0:
                        continue                                   # missing file
                    rng = np.random.default_rng(1000 + k)
                    n = synth.lengths_uniform(rng, 14, 2, 40)      # includes n = 2 (dropped by the loader)
                    pts, off = synth.random_walk_csr(n, 5000 + k)
                    if k % 5 == 0:                                 # a NaN point in the third polyline
                        pts[off[2] + 1, k % 3] = np.nan
                    if k % 7 == 0:                                 # a zero-length polyline
                        pts[off[4]:off[5]] = pts[off[4]]
                    if k == 23:                                    # nothing survives
                        pts, off = synth.lines_to_csr([np.zeros((2, 3)), np.ones((5, 3)), np.full((4, 3), np.nan)])
                    gz = (k % 3 != 0)
                    name = f"{tract}_curves.vtk" + (".gz" if gz else "")
                    vtk_io.write_polylines(os.path.join(d, name), pts, off, binary=(k % 4 != 0),
                                           point_dtype="float" if k % 2 else "double",
                                           layout="offsets" if k % 6 == 0 else "classic")
    with open(os.path.join(root, "tract_config.json"), "w") as f:
        json.dump(CONFIG, f)
    return data, os.path.join(root, "tract_config.json")
