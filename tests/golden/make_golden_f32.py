#!/usr/bin/env python3
"""What a `POINTS n float` tract file means for parity: the UNMODIFIED reference computes
such a file in float32 (pyvista hands it float32 points), this repo upcasts exactly and computes in float64.

    python tests/golden/make_golden_f32.py      # rewrites tests/golden/golden_f32.npz

Stores, for the first 200 polylines of BASELINE configs[0] rounded to float32: `sl32` = the reference's df_sl on the
float32 points (its literal output for such a file) and `sl64` = the reference's df_sl on the same values as float64
(the contract this repo matches to 1e-9)."""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from lesion_condition_vae_b200 import synth  # noqa: E402
from oracle import reference_runner as rr  # noqa: E402

warnings.filterwarnings("ignore")


def main():
    pts, off = synth.config1()
    off = off[:201]
    p32 = pts[:off[-1]].astype(np.float32)
    sl32, _ = rr.reference_compute(p32, off)
    sl64, _ = rr.reference_compute(p32.astype(np.float64), off)
    assert len(sl32) == len(sl64) == 200
    out = os.path.join(HERE, "golden_f32.npz")
    np.savez_compressed(out, points32=p32, offsets=off, sl32=sl32.to_numpy(np.float64), sl64=sl64.to_numpy(np.float64))
    print("wrote", out, os.path.getsize(out))


if __name__ == "__main__":
    main()
