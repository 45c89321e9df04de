"""The batched driver (lesion_condition_vae_b200/tract_driver.py) against golden CSVs produced by the
UNMODIFIED reference driver on the same synthetic study tree (tests/golden/make_driver_golden.py).

CPU part: the host logic (file discovery, gz / ASCII / float parsing, prefix rule, batching, row order,
frame schema) with the device call replaced by an oracle-backed hook.  GPU part: the real thing."""
import numpy as np
import pandas as pd
import pytest

import dataset_fixture
from lesion_condition_vae_b200 import tract_driver as td
from oracle import streamline_oracle as so
from parity_rules import ATOL, BUNDLE_SOURCE, RTOL

GOLDEN = {100: "driver_ms100.csv", 5: "driver_ms5.csv"}
SRC = (0, 2, 4, 6, 7, 8, 10, 11, 12, 16, 13, 14, 15)


def _oracle_compute(points, offsets, bundle_offsets):
    """Stand-in for the device call, same contract as tract_driver._default_compute."""
    B = len(bundle_offsets) - 1
    n_sl = np.zeros(B, np.int64); means = np.full((B, 13), np.nan)
    table, src = so.per_streamline_table(np.asarray(points, dtype=np.float64), offsets)
    for b in range(B):
        rows = table[(src >= bundle_offsets[b]) & (src < bundle_offsets[b + 1])]
        n_sl[b] = len(rows)
        if len(rows):
            with np.errstate(all="ignore"):
                means[b] = np.nanmean(rows[:, SRC], axis=0)
    return n_sl, means


def _bundle_rows(data_root, subject, timepoint, tract, ms):
    """Reference rows (oracle, float64 points as the golden generator fed them) of one study file, for the row-wise tolerance."""
    import glob
    import os
    from lesion_condition_vae_b200 import vtk_io
    hits = glob.glob(os.path.join(data_root, str(subject), str(timepoint), "bundles", f"{tract}_curves.vtk*"))
    assert len(hits) == 1, hits
    pts, off = vtk_io.read_polylines_csr(hits[0])
    sl, _ = so.compute_streamline_metrics_csr(np.asarray(pts, dtype=np.float64), off, max_streamlines=ms)
    return sl.to_numpy()


def _check(df, golden_csv, data_root=None, ms=None):
    import os
    from parity_rules import bundle_tolerances, record
    ref = pd.read_csv(os.path.join(os.path.dirname(__file__), "golden", golden_csv), dtype={"subject_id": str})
    assert list(df.columns) == list(ref.columns)
    assert len(df) == len(ref)
    for c in td.META_COLUMNS:
        assert df[c].astype(str).tolist() == ref[c].astype(str).tolist(), c          # same rows, same order
    assert df["n_streamlines"].dtype == np.float64 and np.array_equal(df["n_streamlines"], ref["n_streamlines"])
    # Tolerance of a bundle mean = the parity rule of its rows, averaged (parity_rules.bundle_tolerances).  This study
    # has polylines of 3-6 points (planar or nearly so: lambda1/lambda3 up to ~1e9), where the rule for the two
    # eigen ratios is |d lambda| <= 1e-12 lambda1 — evaluated row by row from the study's own files.
    worst = {}
    for i in range(len(ref)):
        rows = _bundle_rows(data_root, ref["subject_id"][i], ref["timepoint"][i], ref["tract"][i], ms) if data_root else None
        tols = bundle_tolerances(rows) if rows is not None and len(rows) else None
        for j, (name, src) in enumerate(zip(list(ref.columns)[1:14], BUNDLE_SOURCE)):
            g, r = float(df[name].iloc[i]), float(ref[name].iloc[i])
            if np.isnan(r) or np.isinf(r):
                assert (np.isnan(g) and np.isnan(r)) or g == r, (name, i, g, r)
                continue
            tol = tols[j] if tols is not None else RTOL * abs(r) + ATOL[src]
            worst[name] = max(worst.get(name, 0.0), abs(g - r) / (tol + 1e-300))
            assert abs(g - r) <= tol, (name, i, g, r, tol)
    record(f"reference driver CSV {golden_csv} (bundle means of the synthetic study)", worst)


@pytest.fixture(scope="module")
def study(tmp_path_factory):
    root = tmp_path_factory.mktemp("study")
    data, cfg = dataset_fixture.build(root)
    return root, data, cfg


@pytest.mark.parametrize("ms", [100, 5])
@pytest.mark.parametrize("batch", ["all", "subject", "timepoint"])
def test_driver_host_logic_matches_reference_driver(study, ms, batch):
    root, data, cfg = study
    df = td.process_all_tracts(td.load_config(cfg), data, root / "out", max_streamlines=ms, compute=_oracle_compute, batch=batch)
    _check(df, GOLDEN[ms], data, ms)


def test_select_prefix_rule():
    from lesion_condition_vae_b200 import synth
    pts, off = synth.lines_to_csr(synth.adversarial_lines())
    assert td.select_prefix(pts, off, None).tolist() == [0, 1, 2, 3, 6, 7, 8, 9, 11, 12, 13]   # n=2, NaN, inf lines dropped
    assert td.select_prefix(pts, off, 5).tolist() == [0, 1, 2, 3, 6]
    assert td.select_prefix(pts, off, 0).tolist() == [0]
    p, o = td.gather_polylines(pts, off, np.array([1, 3, 7]))
    assert np.array_equal(p[o[1]:o[2]], pts[off[3]:off[4]]) and o[-1] == len(p)


def test_summary_statistics_and_main(study, monkeypatch):
    root, data, cfg = study
    monkeypatch.setattr(td, "_default_compute", _oracle_compute)
    df = td.main(data_dir=data, output_dir=root / "res", config_path=cfg, max_streamlines=100)
    assert (root / "res" / "comprehensive_tract_geometry_metrics.csv").exists()
    s1 = pd.read_csv(root / "res" / "summary_statistics_by_group_timepoint.csv")
    s2 = pd.read_csv(root / "res" / "summary_statistics_by_tract_group.csv")
    assert len(s1) == 3 * 4 and set(s1["group"]) == {"Sham", "TBI", "PTE"} and len(s2) == 16 * 3
    assert np.isclose(s1.loc[(s1.group == "Sham") & (s1.timepoint == "2d"), "length_mean_mean"].iloc[0],
                      df[(df.group == "Sham") & (df.timepoint == "2d")]["length_mean"].mean())


@pytest.mark.gpu
@pytest.mark.parametrize("ms", [100, 5])
def test_driver_on_gpu_matches_reference_driver(study, ms):
    root, data, cfg = study
    df = td.process_all_tracts(td.load_config(cfg), data, root / "out_gpu", max_streamlines=ms)
    _check(df, GOLDEN[ms], data, ms)
