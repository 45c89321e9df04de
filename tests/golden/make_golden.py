#!/usr/bin/env python3
"""Generate the golden fixtures by running the UNMODIFIED reference in this container.

    python tests/golden/make_golden.py          # rewrites tests/golden/golden_v1.npz

Each case stores its inputs (points float64, offsets int64, max_streamlines) and the reference's
outputs: ``sl`` = df_sl.to_numpy() (rows x 17, reference column order) and ``bundle`` = the 14
df_bundle values (n_streamlines first).  A case whose reference call raises KeyError('length')
(empty result) stores ``sl`` with zero rows and ``bundle`` = all-NaN with count 0.

The reference is imported from /root/reference through oracle/reference_runner.py (stub pyvista,
SURVEY.md §8c); nothing is copied from it.  numpy 2.3.5 / pandas 3.0.2 / Python 3.12.3.
"""
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


def run_case(points, offsets, max_streamlines=None):
    try:
        df_sl, df_b = rr.reference_compute(points, offsets, max_streamlines)
        sl = df_sl.to_numpy(dtype=np.float64).reshape(len(df_sl), 17)
        b = df_b.iloc[0].to_numpy(dtype=np.float64)
        assert list(df_sl.columns) == list(EXPECT_SL), df_sl.columns
        assert list(df_b.columns) == list(EXPECT_B), df_b.columns
    except KeyError as e:  # empty result
        assert e.args == ("length",), e.args
        sl = np.empty((0, 17))
        b = np.full(14, np.nan); b[0] = 0
    return sl, b


from oracle.streamline_oracle import SL_COLUMNS as EXPECT_SL, BUNDLE_COLUMNS  # noqa: E402
EXPECT_B = ("n_streamlines",) + tuple(n for n, _ in BUNDLE_COLUMNS)


def main():
    cases = {}

    def add(name, pts, off, ms=None, input_of=None):
        sl, b = run_case(pts, off, ms)
        cases[name] = dict(max_streamlines=np.int64(-1 if ms is None else ms), sl=sl, bundle=b)
        if input_of is None:
            cases[name].update(points=pts, offsets=off)
        else:  # same input as another case: store it once
            cases[name].update(input_of=np.str_(input_of))
        print(f"{name:28s} S={len(off)-1:6d} P={len(pts):8d} rows={len(sl):6d}")

    # BASELINE.json configs[0]: one 1,000-streamline tract, max_streamlines=1000
    pts, off = synth.config1()
    add("config1", pts, off, 1000)
    add("config1_first100", pts, off, 100, input_of="config1")   # the shipped driver's operating point (max_streamlines=100)

    # adversarial polylines: straight, planar, n=3, n=2, NaN, zero length, duplicates, inf, far offset ...
    pts, off = synth.lines_to_csr(synth.adversarial_lines())
    add("adversarial", pts, off, None)
    for ms in (1, 3, 5, 6, 8):
        add(f"adversarial_max{ms}", pts, off, ms, input_of="adversarial")

    # every length 0..12 plus a few long ones (ragged / tiny inputs)
    rng = np.random.default_rng(99)
    n = np.array(list(range(0, 13)) + [31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 258, 300, 511, 512, 513, 777, 1500], dtype=np.int64)
    pts, off = synth.random_walk_csr(n, 4242)
    add("ragged_small", pts, off, None)

    # heavy-tailed lengths (config 4 law, small sample)
    n = synth.lengths_heavy_tail(rng, 200, 10, 2000)
    pts, off = synth.random_walk_csr(n, 4)
    add("heavy_tail", pts, off, None)

    # nothing survives
    pts, off = synth.lines_to_csr([np.zeros((2, 3)), np.ones((5, 3)), np.full((4, 3), np.nan)])
    add("all_dropped", pts, off, None)
    add("empty", np.empty((0, 3)), np.zeros(1, np.int64), None)

    # four small bundles of config 2 (per-bundle aggregate is the unit the driver consumes)
    for t, k in ((0, 0), (3, 1), (7, 2), (15, 3)):
        pts, off = synth.config2_bundle(t, k, S=60)
        add(f"config2_t{t}_tp{k}", pts, off, None)

    # low-noise curves: large lambda1/lambda3, exercises the conditioned tolerance rule
    n = synth.lengths_uniform(rng, 60, 30, 200)
    pts, off = synth.random_walk_csr(n, 77, sigma=0.004)
    add("low_noise", pts, off, None)

    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat[f"{name}/{k}"] = v
    out = os.path.join(HERE, "golden_v1.npz")
    np.savez_compressed(out, **flat)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
