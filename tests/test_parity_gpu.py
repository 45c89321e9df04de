"""GPU suite: the CUDA path, called through the C ABI (include/tractgeom.h via ctypes), against the
CPU oracle and the reference-generated golden vectors, under the 1e-9 rule of parity_rules.py.

Every test here needs a B200 and libtractgeom.so; a missing library or device FAILS the test."""
import numpy as np
import pytest

from lesion_condition_vae_b200 import _lib, synth, vtk_io
from lesion_condition_vae_b200 import tract_geom_proc as tgp
from oracle import streamline_oracle as so
from parity_rules import ATOL, COLUMNS, RTOL, assert_bundle_close, assert_table_close, record

pytestmark = pytest.mark.gpu

GOLDEN_CASES = [
    "config1", "config1_first100", "adversarial", "adversarial_max1", "adversarial_max3", "adversarial_max5",
    "adversarial_max6", "adversarial_max8", "ragged_small", "heavy_tail",
    "config2_t0_tp0", "config2_t3_tp1", "config2_t7_tp2", "config2_t15_tp3", "low_noise",
]


def _oracle_keep(points, offsets):
    """Expected keep flags, from the oracle's two filters (ref:21 and ref:160)."""
    S = len(offsets) - 1
    keep = np.zeros(S, np.uint8)
    for s in range(S):
        sl = points[offsets[s]:offsets[s + 1]]
        if so.loader_accepts(sl):
            keep[s] |= 1
            if float(np.linalg.norm(np.diff(sl, axis=0), axis=1).sum()) > so.MIN_LEN:
                keep[s] |= 2
    return keep


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_vectors(gpu_ctx, golden, name):
    c = golden[name]
    df_sl, df_b = tgp.compute_streamline_metrics_csr(c["points"], c["offsets"], c["max_streamlines"], ctx=gpu_ctx)
    assert list(df_sl.columns) == list(COLUMNS)
    assert len(df_sl) == len(c["sl"])                      # streamline count: bit-exact
    assert df_sl.index.equals(__import__("pandas").RangeIndex(len(df_sl)))
    assert_table_close(df_sl.to_numpy(), c["sl"], name)
    assert df_b["n_streamlines"].dtype == np.int64 and int(df_b["n_streamlines"].iloc[0]) == int(c["bundle"][0])
    assert_bundle_close(df_b.iloc[0].to_numpy(float), c["bundle"], c["sl"], name)


@pytest.mark.parametrize("name", ["all_dropped", "empty"])
def test_empty_result_raises_keyerror_length(gpu_ctx, golden, name):
    c = golden[name]
    with pytest.raises(KeyError) as e:
        tgp.compute_streamline_metrics_csr(c["points"], c["offsets"], c["max_streamlines"], ctx=gpu_ctx)
    assert e.value.args == ("length",)


def test_keep_flags_and_nan_rows(gpu_ctx, golden):
    c = golden["adversarial"]
    out, keep, sums, counts = gpu_ctx.metrics_host(c["points"], c["offsets"])
    exp = _oracle_keep(c["points"], c["offsets"])
    # the length bit is only defined for loader-accepted rows
    assert np.array_equal(keep & 1, exp & 1)
    assert np.array_equal((keep == 3), (exp == 3))
    assert np.isnan(out[:, keep != 3]).all()
    assert counts[0, 0] == int((exp == 3).sum())


@pytest.mark.parametrize("seed,law", [(21, "uniform"), (22, "normal"), (23, "heavy")])
def test_random_tracts_vs_oracle(gpu_ctx, seed, law):
    rng = np.random.default_rng(seed)
    if law == "uniform":
        n = synth.lengths_uniform(rng, 600, 3, 160)
    elif law == "normal":
        n = synth.lengths_normal(rng, 600)
    else:
        n = synth.lengths_heavy_tail(rng, 400, 10, 5000)
    pts, off = synth.random_walk_csr(n, seed)
    df_sl, df_b = tgp.compute_streamline_metrics_csr(pts, off, ctx=gpu_ctx)
    ref_sl, ref_b = so.compute_streamline_metrics_csr(pts, off)
    assert len(df_sl) == len(ref_sl)
    assert_table_close(df_sl.to_numpy(), ref_sl.to_numpy(), law)
    assert_bundle_close(df_b.iloc[0].to_numpy(float), ref_b.iloc[0].to_numpy(float), ref_sl.to_numpy(), law)


@pytest.mark.parametrize("sigma", [0.2, 0.45, 0.9])
def test_sharp_turns(gpu_ctx, sigma):
    """Turning angles around and beyond the 60-degree limit of the speculative angle series: polylines
    below it stay on the fast path (series + fp32 tail), the others are recomputed exactly; both must
    match the oracle."""
    rng = np.random.default_rng(51)
    pts, off = synth.random_walk_csr(synth.lengths_uniform(rng, 400, 5, 90), 51, sigma=sigma)
    df_sl, _ = tgp.compute_streamline_metrics_csr(pts, off, ctx=gpu_ctx)
    ref_sl, _ = so.compute_streamline_metrics_csr(pts, off)
    assert_table_close(df_sl.to_numpy(), ref_sl.to_numpy(), f"sigma={sigma}")


def test_far_from_origin_covariance(gpu_ctx):
    """a one-pass raw-coordinate covariance fails at +1000 mm; ours must not."""
    rng = np.random.default_rng(31)
    pts, off = synth.random_walk_csr(synth.lengths_uniform(rng, 200, 20, 120), 31)
    pts = pts + np.array([1000.0, -2000.0, 500.0])
    df_sl, _ = tgp.compute_streamline_metrics_csr(pts, off, ctx=gpu_ctx)
    ref_sl, _ = so.compute_streamline_metrics_csr(pts, off)
    assert_table_close(df_sl.to_numpy(), ref_sl.to_numpy(), "far")


def test_float32_points_are_upcast_exactly(gpu_ctx):
    """float32 storage: the device upcasts exactly, so the result equals the float64 run on the
    same (float32-representable) values."""
    pts, off = synth.config1(S=200, seed=8)
    p32 = pts.astype(np.float32)
    a = gpu_ctx.metrics_host(p32, off)[0]
    b = gpu_ctx.metrics_host(p32.astype(np.float64), off)[0]
    assert np.array_equal(a, b, equal_nan=True)
    ref_sl, _ = so.compute_streamline_metrics_csr(p32.astype(np.float64), off)
    assert_table_close(a.T, ref_sl.to_numpy(), "f32")


def test_vtk_file_drop_in(gpu_ctx, tmp_path, golden):
    """The reference-facing call: a path to a legacy VTK file in, two DataFrames out."""
    c = golden["config1"]
    for kw in (dict(binary=True, point_dtype="double"), dict(binary=False, point_dtype="double", layout="offsets")):
        p = vtk_io.write_polylines(tmp_path / "t.vtk", c["points"], c["offsets"], **kw)
        df_sl, df_b = tgp.compute_streamline_metrics(str(p), max_streamlines=1000)
        assert_table_close(df_sl.to_numpy(), c["sl"], "vtk")
        assert_bundle_close(df_b.iloc[0].to_numpy(float), c["bundle"], c["sl"], "vtk")
    df_sl, df_b = tgp.compute_streamline_metrics(str(p), 100)
    c = golden["config1_first100"]
    assert_table_close(df_sl.to_numpy(), c["sl"], "vtk100")
    with pytest.raises(FileNotFoundError):
        tgp.compute_streamline_metrics(str(tmp_path / "nope.vtk"))


def test_batched_bundles_match_per_bundle_calls(gpu_ctx, golden):
    """BASELINE config 2 shape: several bundles, one launch; each bundle equals its own golden case."""
    names = ["config2_t0_tp0", "config2_t3_tp1", "config2_t7_tp2", "config2_t15_tp3"]
    P, O, B, base = [], [np.zeros(1, np.int64)], [0], 0
    for nm in names:
        c = golden[nm]
        P.append(c["points"]); O.append(c["offsets"][1:] + base); base += int(c["offsets"][-1]); B.append(B[-1] + len(c["offsets"]) - 1)
    # an empty bundle in the middle and one that drops everything at the end
    pts = np.concatenate(P + [np.zeros((4, 3))]); off = np.concatenate(O + [np.array([base + 2, base + 4])])
    # B = [0, s1, s2, s3, s4]; duplicate s2 -> empty bundle #2; trailing bundle of two n=2 polylines
    bo = np.array(B[:3] + [B[2]] + B[3:] + [B[-1] + 2], dtype=np.int64)
    res = tgp.compute_bundles_csr(pts, off, bo, ctx=gpu_ctx)
    assert len(res) == 6
    got = [res[0], res[1], res[3], res[4]]
    assert res[2] == (None, None) and res[5] == (None, None)
    for nm, (df_sl, df_b) in zip(names, got):
        c = golden[nm]
        assert_table_close(df_sl.to_numpy(), c["sl"], nm)
        assert_bundle_close(df_b.iloc[0].to_numpy(float), c["bundle"], c["sl"], nm)


def test_opt_in_bundle_spread_columns(gpu_ctx, golden):
    """std / min / max of the 13 bundle columns as OPT-IN extra columns.  The default
    df_bundle keeps the reference's 14 columns; the 39 extras equal np.nanstd / np.nanmin / np.nanmax
    (what ref:193 `_safe_std` defines) over the oracle's df_sl, NaN skipped, inf kept (std NaN then)."""
    import warnings
    from parity_rules import BUNDLE_SOURCE, column_tolerances
    names = ["config2_t0_tp0", "config2_t3_tp1", "config2_t7_tp2"]
    lines = []
    bo = [0]
    for nm in names:
        c = golden[nm]
        lines += [c["points"][c["offsets"][i]:c["offsets"][i + 1]] for i in range(len(c["offsets"]) - 1)]
        bo.append(len(lines))
    # bundle 3: a straight line (elongation / planarity = inf), a planar arc, short and dropped polylines
    t = np.linspace(0.0, 1.0, 30)
    extra = [np.outer(t, [3.0, 4.0, 12.0]), np.stack([np.cos(t), np.sin(t), 0 * t], axis=1) * 20.0,
             np.zeros((2, 3)), synth.random_walk_csr(np.array([40]), seed=77)[0], np.ones((5, 3))]
    lines += extra
    bo += [len(lines), len(lines)]                                   # + an empty bundle
    pts, off = synth.lines_to_csr(lines)
    bo = np.asarray(bo, dtype=np.int64)
    plain = tgp.compute_bundles_csr(pts, off, bo, ctx=gpu_ctx)
    full = tgp.compute_bundles_csr(pts, off, bo, ctx=gpu_ctx, extra_stats=True)
    assert full[4] == (None, None)
    for b in range(4):
        df_sl, df_b = full[b]
        assert list(plain[b][1].columns) == list(tgp.BUNDLE_COLUMNS)                       # default schema untouched
        assert list(df_b.columns) == list(tgp.BUNDLE_COLUMNS) + list(tgp.SPREAD_COLUMNS)
        assert df_b.iloc[0, :14].equals(plain[b][1].iloc[0])
        p, o = synth.lines_to_csr(lines[bo[b]:bo[b + 1]])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref_sl, _ = so.compute_streamline_metrics_csr(p, o)
            for j, src in enumerate(BUNDLE_SOURCE):
                col = ref_sl[src].to_numpy()
                exp = (np.nanstd(col), np.nanmin(col), np.nanmax(col))
                for stat, e in zip(("std", "min", "max"), exp):
                    g = float(df_b[f"{src}_{stat}"].iloc[0])
                    if not np.isfinite(e):
                        assert (np.isnan(e) and np.isnan(g)) or g == e, (b, src, stat, g, e)
                        continue
                    # a statistic of a column is as good as its worst row: the largest row tolerance of the parity rule
                    row_tol = column_tolerances(ref_sl.to_numpy())[:, COLUMNS.index(src)]
                    tol = float(np.max(row_tol[np.isfinite(col)])) + RTOL * abs(e)
                    assert abs(g - e) <= tol, (b, src, stat, g, e, tol)
    # bundle 3 really exercised the special values
    assert np.isinf(full[3][1]["elongation_ratio_max"].iloc[0]) and np.isnan(full[3][1]["elongation_ratio_std"].iloc[0])
    # single-file entry points: reference signature unchanged, extended variant separate
    df_sl, df_b = tgp.compute_streamline_metrics_csr(pts[:off[bo[1]]], off[:bo[1] + 1], ctx=gpu_ctx, extra_stats=True)
    assert df_b.shape == (1, 14 + 39) and df_b.iloc[0].equals(full[0][1].iloc[0])


def test_bundle_spread_device_abi_with_mask(gpu_ctx):
    """tg_bundle_spread_dev on a device-resident table with a selection mask: equals numpy on the selected rows."""
    import torch
    dev = torch.device("cuda:0")
    S = 30_000
    n = synth.torch_lengths("uniform", S, 11, dev)
    pts, off = synth.torch_random_walk_csr(n, 11, dev)
    out = torch.empty((17, S), dtype=torch.float64, device=dev)
    keep = torch.empty(S, dtype=torch.uint8, device=dev)
    gpu_ctx.metrics_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, pts.shape[0], out.data_ptr(), keep.data_ptr())
    sel = (torch.arange(S, device=dev) % 3 != 0).to(torch.uint8)
    bo = np.array([0, 9_000, 9_000, 21_500, S], dtype=np.int64)
    B = len(bo) - 1
    sums = torch.empty((B, 13), dtype=torch.float64, device=dev)
    counts = torch.empty((B, 14), dtype=torch.int64, device=dev)
    spread = torch.empty((B, 13, 3), dtype=torch.float64, device=dev)
    gpu_ctx.bundle_reduce_dev(out.data_ptr(), keep.data_ptr(), sel.data_ptr(), S, bo, sums.data_ptr(), counts.data_ptr())
    gpu_ctx.bundle_spread_dev(out.data_ptr(), keep.data_ptr(), sel.data_ptr(), S, bo, sums.data_ptr(), counts.data_ptr(), spread.data_ptr())
    gpu_ctx.synchronize()
    table, mask, got = out.cpu().numpy(), sel.cpu().numpy().astype(bool), spread.cpu().numpy()
    src = (0, 2, 4, 6, 7, 8, 10, 11, 12, 16, 13, 14, 15)
    for b in range(B):
        rows = table[:, bo[b]:bo[b + 1]][:, mask[bo[b]:bo[b + 1]]]
        if rows.shape[1] == 0:
            assert np.isnan(got[b]).all()
            continue
        for j, m in enumerate(src):
            col = rows[m]
            np.testing.assert_allclose(got[b, j], [np.std(col), col.min(), col.max()], rtol=1e-12, atol=1e-15)


def test_device_pointer_abi_and_size_independent_properties(gpu_ctx):
    """Device-resident call at a size the oracle cannot cover, checked through invariances:
    rigid translation leaves every metric but the centroid unchanged (to rounding), reversing each
    polyline preserves length / chord / bbox / centroid / eigen ratios, and a subsample matches the oracle."""
    import torch
    dev = torch.device("cuda:0")
    S = 200_000
    n = synth.torch_lengths("normal", S, 3, dev)
    pts, off = synth.torch_random_walk_csr(n, 3, dev)
    P = pts.shape[0]
    assert int(off[-1]) == P

    def run(p):
        out = torch.empty((17, S), dtype=torch.float64, device=dev)
        keep = torch.empty(S, dtype=torch.uint8, device=dev)
        gpu_ctx.metrics_dev(p.data_ptr(), _lib.F64, off.data_ptr(), S, P, out.data_ptr(), keep.data_ptr())
        gpu_ctx.synchronize()
        return out, keep

    torch.cuda.synchronize()
    out, keep = run(pts)
    assert int((keep == 3).sum()) == S                     # count bit-exact: every polyline survives
    # subsample vs oracle
    idx = np.unique(np.concatenate([np.arange(300), np.random.default_rng(0).integers(0, S, 300),
                                    torch.topk(n, 20).indices.cpu().numpy()]))
    off_h = off.cpu().numpy()
    rows = []
    for s in idx:
        sl = pts[off_h[s]:off_h[s + 1]].cpu().numpy()
        rows.append(so.metrics_row(sl))
    assert_table_close(out[:, torch.as_tensor(idx, device=dev)].T.cpu().numpy(), np.asarray(rows), "subsample")
    # translation
    shift = torch.tensor([3.0, -7.0, 11.0], dtype=torch.float64, device=dev)
    out_t, _ = run(pts + shift)
    same = [m for m in range(17) if m not in (13, 14, 15)]
    a, b = out[same].cpu().numpy(), out_t[same].cpu().numpy()
    assert np.all(np.abs(a - b) <= 1e-7 * np.abs(a) + 1e-9)
    assert torch.allclose(out_t[13:16], out[13:16] + shift[:, None], rtol=0, atol=1e-9)
    # bundle reduce on device vs torch
    sums = torch.empty((1, 13), dtype=torch.float64, device=dev)
    counts = torch.empty((1, 14), dtype=torch.int64, device=dev)
    gpu_ctx.bundle_reduce_dev(out.data_ptr(), keep.data_ptr(), 0, S, np.array([0, S]), sums.data_ptr(), counts.data_ptr())
    gpu_ctx.synchronize()
    src = torch.tensor([0, 2, 4, 6, 7, 8, 10, 11, 12, 16, 13, 14, 15], device=dev)
    exp = out[src].sum(dim=1)
    assert torch.allclose(sums[0], exp, rtol=1e-11, atol=0)
    assert counts[0].tolist() == [S] * 14
    assert gpu_ctx.launches > 0


def test_chunked_host_pipeline_equals_single_shot(gpu_ctx, monkeypatch):
    """tg_metrics_csr_host streams the tractogram in polyline-aligned chunks (H2D / kernels / D2H
    overlapped); the table must be bit-identical whatever the chunk size, for float64 and float32."""
    rng = np.random.default_rng(41)
    n = np.concatenate([synth.lengths_uniform(rng, 700, 0, 90), [2500, 3, 2047, 2046, 4100]])
    pts, off = synth.random_walk_csr(n, 41)
    bo = np.array([0, 100, 100, 650, len(n)], dtype=np.int64)
    for arr in (pts, pts.astype(np.float32)):
        monkeypatch.delenv("TG_HOST_CHUNK_POINTS", raising=False)
        a = gpu_ctx.metrics_host(arr, off, bo)
        for cp in ("1", "777", "20000"):
            monkeypatch.setenv("TG_HOST_CHUNK_POINTS", cp)
            b = gpu_ctx.metrics_host(arr, off, bo)
            assert np.array_equal(a[0], b[0], equal_nan=True) and np.array_equal(a[1], b[1])
            assert np.array_equal(a[3], b[3]) and np.allclose(a[2], b[2], rtol=1e-13, atol=0)
    monkeypatch.delenv("TG_HOST_CHUNK_POINTS", raising=False)
    ref_sl, _ = so.compute_streamline_metrics_csr(pts, off)
    out, keep, _, _ = gpu_ctx.metrics_host(pts, off)
    assert_table_close(out[:, keep == 3].T, ref_sl.to_numpy(), "chunked+long")


def test_config2_batch_of_64_bundles(gpu_ctx):
    """BASELINE configs[1]: 16 tracts x 4 timepoints, ~5k polylines each, ONE batched call.
    Every bundle's n_streamlines is exact; three bundles are checked row by row against the oracle."""
    pts, off, bo = synth.config2(S=5000)
    assert len(bo) == 65 and len(off) - 1 == 320_000
    res = tgp.compute_bundles_csr(pts, off, bo, ctx=gpu_ctx)
    assert len(res) == 64
    for b, (df_sl, df_b) in enumerate(res):
        assert len(df_sl) == 5000 and int(df_b["n_streamlines"].iloc[0]) == 5000
    for b in (0, 37, 63):
        lo, hi = int(bo[b]), int(bo[b + 1])
        p, o = pts[off[lo]:off[hi]], off[lo:hi + 1] - off[lo]
        ref_sl, ref_b = so.compute_streamline_metrics_csr(p[:int(o[600])], o[:601])     # first 600 rows of the bundle
        assert_table_close(res[b][0].to_numpy()[:600], ref_sl.to_numpy(), f"bundle{b}")
        # bundle means: against the oracle's nan-mean over OUR full table (the aggregate itself is exact arithmetic)
        mine = res[b][0]
        exp = so.bundle_summary(mine).iloc[0].to_numpy(float)
        assert_bundle_close(res[b][1].iloc[0].to_numpy(float), exp, mine.to_numpy(), f"bundle{b} means")


def test_heavy_tail_at_scale_subsample(gpu_ctx):
    """BASELINE configs[3] law (n = min(5000, floor(10/U))) at 300k polylines on the device: counts exact,
    the 60 longest polylines (long-polyline kernel, n > 2046) and a random subsample match the oracle."""
    import torch
    dev = torch.device("cuda:0")
    S = 300_000
    n = synth.torch_lengths("heavy", S, 4, dev)
    pts, off = synth.torch_random_walk_csr(n, 4, dev)
    P = pts.shape[0]
    out = torch.empty((17, S), dtype=torch.float64, device=dev)
    keep = torch.empty(S, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.metrics_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, out.data_ptr(), keep.data_ptr())
    gpu_ctx.synchronize()
    assert int((keep == 3).sum()) == S and int(n.max()) == 5000 and int((n > 2046).sum()) > 100
    idx = np.unique(np.concatenate([torch.topk(n, 60).indices.cpu().numpy(), np.random.default_rng(1).integers(0, S, 400)]))
    off_h = off.cpu().numpy()
    rows = [so.metrics_row(pts[off_h[s]:off_h[s + 1]].cpu().numpy()) for s in idx]
    assert_table_close(out[:, torch.as_tensor(idx, device=dev)].T.cpu().numpy(), np.asarray(rows), "heavy subsample")


def test_heavy_tail_long_row_of_the_queue(gpu_ctx):
    """Same law at 800k polylines: more than 6144 polylines exceed 1024 points, so they take row 0 of the
    queue (groups of nearly equal length, longest first) instead of the long-polyline kernel; two
    polylines beyond the queue's 9216-point limit still go to that kernel.  Counts exact, the longest
    polylines and a random subsample match the oracle, and the rows agree bit for bit with the SAME
    polylines computed in a small table (where they take the long-polyline kernel or other groups)."""
    import torch
    dev = torch.device("cuda:0")
    S = 800_000
    n = synth.torch_lengths("heavy", S, 7, dev)
    n[12345] = 9216; n[23456] = 9217; n[34567] = 12000; n[45678] = 1025; n[56789] = 1024
    assert int((n > 1024).sum()) >= 6144
    pts, off = synth.torch_random_walk_csr(n, 7, dev)
    P = pts.shape[0]
    out = torch.empty((17, S), dtype=torch.float64, device=dev)
    keep = torch.empty(S, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.metrics_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, out.data_ptr(), keep.data_ptr())
    gpu_ctx.synchronize()
    assert int((keep == 3).sum()) == S
    special = np.array([12345, 23456, 34567, 45678, 56789])
    idx = np.unique(np.concatenate([torch.topk(n, 40).indices.cpu().numpy(), special,
                                    torch.nonzero((n > 1024) & (n < 5000)).flatten()[:120].cpu().numpy(),
                                    np.random.default_rng(2).integers(0, S, 300)]))
    off_h = off.cpu().numpy()
    lines = [pts[off_h[s]:off_h[s + 1]].cpu().numpy() for s in idx]
    got = out[:, torch.as_tensor(idx, device=dev)].T.cpu().numpy()
    assert_table_close(got, np.asarray([so.metrics_row(l) for l in lines]), "heavy long-row subsample")
    sub_pts, sub_off = synth.lines_to_csr(lines)                 # few long polylines here: long-polyline kernel
    table, sub_keep = gpu_ctx.metrics_host(sub_pts, sub_off)[:2]
    assert np.all(sub_keep == 3)
    mism = np.nonzero(~np.all((got == table.T) | (np.isnan(got) & np.isnan(table.T)), axis=1))[0]
    # the long-polyline kernel splits a polyline over 32 lanes and merges: same values to rounding, not bit-identical
    short = np.array([len(l) <= 1024 for l in lines])
    assert not np.any(short[mism]), "queued polylines must not depend on the table they are in"
    assert_table_close(got, table.T, "long row vs long-polyline kernel")


FULL_SIZE = [("normal", 10_000_000, 5), ("heavy", 2_000_000, 4), ("normal", 1_000_000, 3)]   # BASELINE configs[4], [3], [2]


def _rule(case, column, got, ref, tol=None):
    """Every polyline of a full-size table against a torch re-derivation: worst |got - ref| / tolerance, recorded in the
    parity budget and asserted <= 1.  Default tolerance = the parity rule RTOL |ref| + ATOL[column]."""
    if tol is None:
        tol = RTOL * ref.abs() + ATOL[column]
    worst = float(((got - ref).abs() / (tol + 1e-300)).max())
    record(case, {column: worst})
    assert worst <= 1.0, f"{case}: {column} error/tolerance = {worst}"


@pytest.mark.parametrize("law,S,seed", FULL_SIZE)
def test_full_size_configs(gpu_ctx, law, S, seed):
    """BASELINE configs[4] (1e7 polylines, ~1e9 points, 24 GB), configs[3] (2e6 polylines, heavy-tailed
    lengths 10..5000) and configs[2] (1e6 polylines) at their full sizes: counts exact; length, chord, tortuosity, straightness, bending
    angle, bounding box, angular dispersion and
    centroid columns of ALL polylines against torch.segment_reduce (independent of the oracle); the bundle
    means against the column means; first / random / longest polylines against the oracle on all 17 columns."""
    import torch
    dev = torch.device("cuda:0")
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~70 GB of device memory")
    n = synth.torch_lengths(law, S, seed, dev)
    pts, off = synth.torch_random_walk_csr(n, seed, dev)
    P = pts.shape[0]
    out = torch.empty((17, S), dtype=torch.float64, device=dev)
    keep = torch.empty(S, dtype=torch.uint8, device=dev)
    sums = torch.empty((1, 13), dtype=torch.float64, device=dev)
    counts = torch.empty((1, 14), dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.metrics_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, out.data_ptr(), keep.data_ptr())
    gpu_ctx.bundle_reduce_dev(out.data_ptr(), keep.data_ptr(), 0, S, np.array([0, S]), sums.data_ptr(), counts.data_ptr())
    gpu_ctx.synchronize()
    case = f"{law} {S} polylines, all rows vs torch re-derivation"
    assert int((keep == 3).sum()) == S and int(counts[0, 0]) == S and int(off[-1]) == P      # counts bit-exact
    assert bool((counts[0, 1:] == S).all())
    # length of every polyline: torch segmented sum of the segment norms (zero at the joints between polylines)
    seg = torch.zeros(P, dtype=torch.float64, device=dev)
    seg[:-1] = torch.linalg.norm(pts[1:] - pts[:-1], dim=1)
    seg[off[1:] - 1] = 0.0
    L = torch.segment_reduce(seg, "sum", lengths=n)
    del seg
    _rule(case, "length", out[0], L)
    # centroid of every polyline, one coordinate at a time
    for c in range(3):
        cen = torch.segment_reduce(pts[:, c].contiguous(), "sum", lengths=n) / n
        _rule(case, COLUMNS[13 + c], out[13 + c], cen)
    # chord, tortuosity, straightness, bounding box of every polyline (ref:35-46, 114-117)
    first, last = pts[off[:-1]], pts[off[1:] - 1]
    chord = torch.sqrt(((last - first) ** 2).sum(dim=1))
    _rule(case, "end_to_end", out[1], chord)
    _rule(case, "tortuosity", out[2], L / chord.clamp_min(1e-8))
    _rule(case, "straightness", out[3], chord / L.clamp_min(1e-8))
    vol = torch.ones(S, dtype=torch.float64, device=dev)
    for c in range(3):
        col = pts[:, c].contiguous()
        vol *= torch.segment_reduce(col, "max", lengths=n) - torch.segment_reduce(col, "min", lengths=n)
        del col
    _rule(case, "bbox_vol", out[9], vol)
    del vol
    # bending angle and angular dispersion of every polyline (ref:98-106, 143-148), re-derived with torch
    d = pts[1:] - pts[:-1]
    t = d / (torch.linalg.norm(d, dim=1, keepdim=True) + 1e-12)
    del d
    ang = torch.zeros(P, dtype=torch.float64, device=dev)
    ang[:-2] = torch.acos(torch.clamp((t[:-1] * t[1:]).sum(dim=1), -1.0, 1.0))
    joint = torch.zeros(P, dtype=torch.bool, device=dev)        # angle i uses points i, i+1, i+2: all three in one polyline
    joint[off[1:] - 1] = True
    joint[(off[1:] - 2).clamp_min(0)] = True
    ang[joint] = 0.0
    bend = torch.segment_reduce(ang, "sum", lengths=n) / (n - 2)
    del ang
    _rule(case, "bend_angle_mean", out[8], bend)
    tsq = torch.zeros(P, dtype=torch.float64, device=dev)
    tsq[:-1] = (t * t).sum(dim=1)
    tsq[off[1:] - 1] = 0.0
    disp = torch.segment_reduce(tsq, "sum", lengths=n) / (n - 1)
    del tsq
    for c in range(3):
        tc = torch.zeros(P, dtype=torch.float64, device=dev)
        tc[:-1] = t[:, c]
        tc[off[1:] - 1] = 0.0
        disp -= (torch.segment_reduce(tc, "sum", lengths=n) / (n - 1)) ** 2
        del tc
    del t, joint
    # mean |t - tbar|^2 = mean |t|^2 - |tbar|^2: the re-derivation cancels (1 - 0.9..), its own rounding is ~1e-15 absolute
    _rule(case, "ang_dispersion", out[16], disp, RTOL * disp.abs() + max(ATOL["ang_dispersion"], 1e-14))
    # bundle means = column means (deterministic tree on the device vs torch)
    src = [0, 2, 4, 6, 7, 8, 10, 11, 12, 16, 13, 14, 15]
    means = (sums[0] / counts[0, 1:]).cpu().numpy()
    ref_means = out[src].mean(dim=1).cpu().numpy()
    np.testing.assert_allclose(means, ref_means, rtol=1e-11, atol=1e-13)
    # all 17 columns of a subsample against the oracle
    idx = np.unique(np.concatenate([np.arange(300), np.random.default_rng(seed).integers(0, S, 300),
                                    torch.topk(n, 20).indices.cpu().numpy(), torch.topk(-n, 20).indices.cpu().numpy()]))
    off_h = off.cpu().numpy()
    rows = [so.metrics_row(pts[off_h[s]:off_h[s + 1]].cpu().numpy()) for s in idx]
    assert_table_close(out[:, torch.as_tensor(idx, device=dev)].T.cpu().numpy(), np.asarray(rows), f"{law} full-size subsample")


@pytest.mark.parametrize("law,S,seed", FULL_SIZE)
def test_full_size_curvature_torsion_eigen_columns(gpu_ctx, law, S, seed):
    """The differential and spectral columns of EVERY polyline at the full BASELINE sizes, re-derived with torch
    from the reference formulas (np.gradient with one-sided ends twice, cross product, ref:48-96; covariance
    eigenvalues, ref:119-141) — no oracle involved."""
    import torch
    dev = torch.device("cuda:0")
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs ~130 GB of device memory")
    n = synth.torch_lengths(law, S, seed, dev)
    pts, off = synth.torch_random_walk_csr(n, seed, dev)
    P = pts.shape[0]
    out = torch.empty((17, S), dtype=torch.float64, device=dev)
    keep = torch.empty(S, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    gpu_ctx.metrics_dev(pts.data_ptr(), _lib.F64, off.data_ptr(), S, P, out.data_ptr(), keep.data_ptr())
    gpu_ctx.synchronize()
    assert int((keep == 3).sum()) == S
    nf = n.to(torch.float64)
    case = f"{law} {S} polylines, all rows vs torch re-derivation"

    # np.gradient along each polyline: (f[next] - f[prev]) * s, next/prev clamped to the polyline, s = 1 at its ends
    i = torch.arange(P, device=dev)
    lo = torch.repeat_interleave(off[:-1], n)
    hi = torch.repeat_interleave(off[1:] - 1, n)
    prv = torch.maximum(i - 1, lo)
    nxt = torch.minimum(i + 1, hi)
    sc = torch.where((i == lo) | (i == hi), 1.0, 0.5).to(torch.float64)[:, None]
    interior = i < hi                                             # j < n-1 (ref:82: left-point rule of the energy)
    del i, lo, hi

    def grad(f):
        return (f[nxt] - f[prv]) * sc

    v = grad(pts)
    a = grad(v)
    b = torch.linalg.cross(v, a)
    del a
    kappa = torch.linalg.norm(b, dim=1) / (torch.linalg.norm(v, dim=1) + 1e-12) ** 3          # ref:57-59
    kmean = torch.segment_reduce(kappa, "sum", lengths=n) / nf
    _rule(case, "curv_mean", out[4], kmean)
    kvar = torch.segment_reduce((kappa - torch.repeat_interleave(kmean, n)) ** 2, "sum", lengths=n) / nf
    _rule(case, "curv_std", out[5], kvar.sqrt())
    ds = torch.zeros(P, dtype=torch.float64, device=dev)
    ds[:-1] = torch.linalg.norm(pts[1:] - pts[:-1], dim=1) + 1e-12                            # ref:77
    energy = torch.segment_reduce(torch.where(interior, kappa * kappa * ds, 0.0), "sum", lengths=n)
    _rule(case, "curv_energy", out[6], energy)
    del kappa, ds, v, interior, kvar
    db = grad(b)                                                                             # ref:91
    tau = (b * db).sum(dim=1) / (torch.linalg.norm(b, dim=1) ** 2 + 1e-12)                    # ref:92-94
    del b, db
    tmean = torch.where(n >= 4, torch.segment_reduce(tau, "sum", lengths=n) / nf, 0.0)        # ref:86,96
    # torsion is a sum that cancels, and this float64 re-derivation carries the rounding of its own terms: against it
    # the rule is 1e-9 of the mean |tau| plus ATOL (the strict rule on |mean tau| is applied to the oracle subsample)
    tabs = torch.segment_reduce(tau.abs(), "sum", lengths=n) / nf
    _rule(case, "torsion_mean (vs mean |tau|)", out[7], tmean, RTOL * tabs + ATOL["torsion_mean"])
    del tau, tabs, prv, nxt, sc

    # covariance eigenvalues (ddof = 1) about the centroid
    C = torch.empty((S, 3, 3), dtype=torch.float64, device=dev)
    cen = [torch.segment_reduce(pts[:, c].contiguous(), "sum", lengths=n) / nf for c in range(3)]
    q = [pts[:, c] - torch.repeat_interleave(cen[c], n) for c in range(3)]
    for r in range(3):
        for c in range(r, 3):
            C[:, r, c] = C[:, c, r] = torch.segment_reduce(q[r] * q[c], "sum", lengths=n) / (nf - 1)
    del q
    lam = torch.linalg.eigvalsh(C.cpu()).flip(1).to(dev)             # LAPACK on the host (descending): cuSOLVER's batched syev rejects this batch
    l1, l2, l3 = lam[:, 0], lam[:, 1], lam[:, 2]
    cond = l1 / l3
    ok = l3 > 1e-12
    assert bool(ok.all())                                                                    # no inf ratios in these laws
    # parity_rules / SURVEY.md N7: relative rule up to l1/l3 = 1e5, |d lambda| <= 1e-12 l1 beyond
    from parity_rules import COND_STRICT, EIG_ATOL
    e, p = l1 / l2, l2 / l3
    strict = cond <= COND_STRICT
    record(case, {"max l1/l3": float(cond.max()), "share of rows beyond l1/l3 = 1e5": float((~strict).double().mean())})
    _rule(case, "elongation_ratio", out[10], e, torch.where(strict, RTOL * e, EIG_ATOL * e * e))
    _rule(case, "planarity_ratio", out[11], p, torch.where(strict, RTOL * p, EIG_ATOL * p * (cond + e)))
    _rule(case, "anisotropy_ratio", out[12], l1 / (l1 + l2 + l3 + 1e-12))


def test_degenerate_grid_polylines(gpu_ctx):
    """Integer-grid polylines (duplicate points, collinear triples, right angles, reversals): almost every
    one leaves the speculative path and is recomputed by the exact pipeline; inf / NaN-to-number / zero
    semantics of the reference (ref:60, 81, 95, 128, 134) must come out the same."""
    import warnings
    rng = np.random.default_rng(2024)
    lines = [rng.integers(-2, 3, size=(int(rng.integers(3, 13)), 3)).astype(np.float64) for _ in range(400)]
    pts, off = synth.lines_to_csr(lines)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_sl, ref_b = so.compute_streamline_metrics_csr(pts, off)
    df_sl, df_b = tgp.compute_streamline_metrics_csr(pts, off, ctx=gpu_ctx)
    assert len(df_sl) == len(ref_sl)
    assert_table_close(df_sl.to_numpy(), ref_sl.to_numpy(), "grid")


@pytest.mark.parametrize("scale", [1e-6, 1e-3, 1e3, 1e6])
def test_coordinate_units(gpu_ctx, scale):
    """The additive 1e-12 / 1e-8 epsilons of the reference make the metrics unit-dependent; the speculative
    path has to notice when its first-order treatment of them stops being valid (tiny units) and hand
    over to the exact pipeline.  Same tractogram in micrometres ... kilometres-ish."""
    import warnings
    rng = np.random.default_rng(61)
    pts, off = synth.random_walk_csr(synth.lengths_uniform(rng, 300, 3, 80), 61)
    pts = pts * scale
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref_sl, _ = so.compute_streamline_metrics_csr(pts, off)
    df_sl, _ = tgp.compute_streamline_metrics_csr(pts, off, ctx=gpu_ctx)
    assert len(df_sl) == len(ref_sl)
    got, ref = df_sl.to_numpy(), ref_sl.to_numpy()
    # absolute floors of parity_rules.py are written for millimetre data: scale them with the unit of each column
    from parity_rules import ATOL, COLUMNS, column_errors
    unit = {"length": 1, "end_to_end": 1, "curv_mean": -1, "curv_std": -1, "curv_energy": -1, "bbox_vol": 3,
            "centroid_x": 1, "centroid_y": 1, "centroid_z": 1}
    saved = dict(ATOL)
    try:
        for k, p in unit.items():
            ATOL[k] = saved[k] * max(scale ** p, 1.0)
        errs = column_errors(got, ref)
    finally:
        ATOL.update(saved)
    bad = {k: v for k, v in errs.items() if not v <= 1.0}
    assert not bad, (scale, bad)


def test_big_endian_storage_and_batch_api_are_bit_identical(gpu_ctx, tmp_path):
    """The file's own storage (big-endian float / double, decoded on the device) gives the same bits as native arrays; a batch
    pushed file by file (tg_batch_push / tg_batch_run, mixed storage types, pinned arena, capacity growth) gives the same
    table and bundle sums as one tg_metrics_csr_host call on the concatenated tractogram."""
    rng = np.random.default_rng(77)
    files = []
    for k in range(7):
        n = synth.lengths_uniform(rng, 300 + 211 * k, 2, 130)
        pts, off = synth.random_walk_csr(n, 500 + k)
        if k == 3:
            pts[off[5] + 1, 2] = np.nan
        files.append((pts.astype(np.float32) if k % 2 else pts, off))
    arena = _lib.PinnedArena(1 << 16)                          # small: must grow
    gpu_ctx.batch_begin(1000, 10)                              # capacities are hints: must grow
    held, cat_p, cat_o, bo = [], [], [np.zeros(1, np.int64)], [0]
    base = 0
    for k, (pts, off) in enumerate(files):
        be = pts.astype(pts.dtype.newbyteorder(">")) if k % 3 else pts
        pinned = arena.take(be.nbytes).view(be.dtype).reshape(be.shape)
        np.copyto(pinned, be)
        held.append(gpu_ctx.batch_push(pinned, off))
        native64 = pts.astype(np.float64)
        o1 = gpu_ctx.metrics_host(native64, off)
        o2 = gpu_ctx.metrics_host(be, off)
        assert np.array_equal(o1[0].view(np.uint64), o2[0].view(np.uint64)) and np.array_equal(o1[1], o2[1])
        cat_p.append(native64); cat_o.append(off[1:] + base); base += int(off[-1]); bo.append(bo[-1] + len(off) - 1)
    out, keep, sums, counts = gpu_ctx.batch_run(np.asarray(bo), want_rows=True)
    ref = gpu_ctx.metrics_host(np.concatenate(cat_p), np.concatenate(cat_o), np.asarray(bo))
    assert np.array_equal(out.view(np.uint64), ref[0].view(np.uint64)) and np.array_equal(keep, ref[1])
    assert np.array_equal(counts, ref[3]) and np.allclose(sums, ref[2], rtol=1e-13, atol=0)
    assert (keep != 3).any()
    arena.close()


def test_compute_files_streams_a_study_like_the_per_file_calls(gpu_ctx, tmp_path):
    """tract_driver.compute_files (parse into pinned memory -> push -> one run) against compute_streamline_metrics file by file,
    with and without the max_streamlines prefix rule, across binary/ASCII, float/double, gz, and an empty result."""
    from lesion_condition_vae_b200 import tract_driver as td
    rng = np.random.default_rng(5)
    paths = []
    for k in range(9):
        pts, off = synth.random_walk_csr(synth.lengths_uniform(rng, 120 + 40 * k, 2, 60), 900 + k)
        if k == 4:
            pts, off = synth.lines_to_csr([np.zeros((2, 3)), np.ones((5, 3))])          # nothing survives
        if k == 6:
            pts[off[1] + 1, 0] = np.inf
        paths.append(vtk_io.write_polylines(tmp_path / f"t{k}.vtk{'.gz' if k % 3 == 0 else ''}", pts, off, binary=(k % 4 != 1),
                                            point_dtype="float" if k % 2 else "double", layout="offsets" if k == 5 else "classic"))
    paths.append(str(tmp_path / "missing.vtk"))
    for ms in (None, 25):
        errors = {}
        n_sl, means = td.compute_files(paths, ms, ctx=gpu_ctx, on_error=lambda i, e: errors.__setitem__(i, e))
        assert set(errors) == {len(paths) - 1} and n_sl[-1] == 0 and n_sl[4] == 0
        for i, p in enumerate(paths[:-1]):
            if i == 4:
                continue
            _, df_b = tgp.compute_streamline_metrics(p, ms)
            row = df_b.iloc[0].to_numpy(float)
            assert n_sl[i] == row[0]
            assert np.allclose(means[i], row[1:], rtol=1e-13, atol=0, equal_nan=True), (i, ms)


def test_float32_file_matches_the_float64_contract_not_the_float32_reference(gpu_ctx):
    """`POINTS n float` data (golden_f32.npz, produced by the unmodified reference): the CUDA path fed the float32 array equals
    the reference run on the exactly upcast values under the 1e-9 rule; against the reference's literal float32 output it
    differs by float32 rounding (1e-7 typical, up to ~1e-4 on torsion / bending angle) — the documented deviation."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_f32.npz"))
    p32, off = z["points32"], z["offsets"]
    assert p32.dtype == np.float32
    out, keep = gpu_ctx.metrics_host(p32, off)[:2]
    assert np.all(keep == 3)
    assert_table_close(out.T, z["sl64"], "float32 file vs float64 contract")
    with np.errstate(all="ignore"):
        rel = np.abs(out.T - z["sl32"]) / np.maximum(np.abs(z["sl32"]), 1e-300)
    assert 1e-8 < np.nanmax(rel[:, 0]) < 1e-6 and np.nanmax(rel) < 1e-3
