"""The parity rule shared by every comparison against the oracle (BASELINE.json: 1e-9 relative).

    |got - ref| <= RTOL * |ref| + ATOL[metric]        inf must equal inf, NaN must equal NaN

RTOL = 1e-9 is BASELINE.json's contract.  ATOL exists only where the reference output itself is rounding noise:
a straight line's curv_mean is ~1e-16 and its ang_dispersion ~2e-31; torsion_mean and the centroid are sums that
cancel; a straight line's bending angle is arccos(1 - 4e-12 +- 1e-16) = 2.8e-6 +- 4e-11, hence 1e-10 for that
column (on ordinary polylines, angle ~0.04 rad, that is 2.5e-9 relative: the relative term decides there).

The two eigenvalue ratios follow SURVEY.md N7 literally: the relative rule while lambda1/lambda3 <= 1e5; beyond,
the eigenvalues themselves must agree to |d lambda_k| <= 1e-12 lambda1 (LAPACK's own error is ~1e-16 lambda1, so
the reference's ratios are good to ~1e-16 lambda1/lambda3 relative and no tighter).  Expressed on the columns the
table holds (e = lambda1/lambda2, p = lambda2/lambda3, lambda1/lambda3 = e p):
    d e <= 1e-12 e^2            (<=> |d (lambda2/lambda1)| <= 1e-12)
    d p <= 1e-12 p (e p + e)    (<=> |d lambda3| and |d lambda2| <= 1e-12 lambda1)

Every comparison also records its worst error/tolerance per column in BUDGET; the GPU test session dumps it to
gpurun_out/parity_budget.json (tests/conftest.py), the source of profiles/parity_budget_r2.json.
"""
import numpy as np

RTOL = 1e-9
COND_STRICT = 1e5          # lambda1/lambda3 up to which the eigen ratios obey the relative rule
EIG_ATOL = 1e-12           # beyond: |d lambda_k| <= EIG_ATOL * lambda1
COLUMNS = (
    "length", "end_to_end", "tortuosity", "straightness", "curv_mean", "curv_std", "curv_energy",
    "torsion_mean", "bend_angle_mean", "bbox_vol", "elongation_ratio", "planarity_ratio",
    "anisotropy_ratio", "centroid_x", "centroid_y", "centroid_z", "ang_dispersion",
)
ATOL = {
    "length": 0.0, "end_to_end": 1e-15, "tortuosity": 0.0, "straightness": 0.0,
    "curv_mean": 1e-13, "curv_std": 1e-13, "curv_energy": 1e-13, "torsion_mean": 1e-12,
    "bend_angle_mean": 1e-10, "bbox_vol": 1e-12, "elongation_ratio": 0.0, "planarity_ratio": 0.0,
    "anisotropy_ratio": 0.0, "centroid_x": 1e-12, "centroid_y": 1e-12, "centroid_z": 1e-12,
    "ang_dispersion": 1e-14,
}
BUNDLE_SOURCE = ("length", "tortuosity", "curv_mean", "curv_energy", "torsion_mean", "bend_angle_mean",
                 "elongation_ratio", "planarity_ratio", "anisotropy_ratio", "ang_dispersion",
                 "centroid_x", "centroid_y", "centroid_z")

BUDGET = {}                # case -> {column: worst error / tolerance seen}


def record(case, errs):
    """Merge per-column error/tolerance ratios into BUDGET[case] (keeping the worst)."""
    slot = BUDGET.setdefault(str(case), {})
    for k, v in errs.items():
        v = float(v)
        if not (slot.get(k, 0.0) >= v):
            slot[k] = v


def eigen_ratio_tolerances(ref):
    """Per-row absolute tolerances of columns 10 (elongation) and 11 (planarity) of a reference table (R,17)."""
    ref = np.asarray(ref, dtype=np.float64)
    e, p = ref[:, 10], ref[:, 11]
    with np.errstate(all="ignore"):
        cond = e * p
        strict = ~(cond > COND_STRICT)              # NaN / inf condition numbers fall under the special-value rule anyway
        tol_e = np.where(strict, RTOL * np.abs(e), EIG_ATOL * e * e)
        tol_p = np.where(strict, RTOL * np.abs(p), EIG_ATOL * np.abs(p) * (cond + np.abs(e)))
    return tol_e, tol_p


def column_tolerances(ref):
    """(R,17) absolute tolerance of every entry of a reference table."""
    ref = np.asarray(ref, dtype=np.float64)
    with np.errstate(all="ignore"):
        tol = RTOL * np.abs(ref) + np.asarray([ATOL[c] for c in COLUMNS])[None, :]
    tol[:, 10], tol[:, 11] = eigen_ratio_tolerances(ref)
    return tol


def column_errors(got, ref):
    """Per-column max of |got-ref| / tolerance over rows (<= 1 passes); got/ref (R,17)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    tol = column_tolerances(ref)
    res = {}
    for m, name in enumerate(COLUMNS):
        g, r = got[:, m], ref[:, m]
        same_special = (np.isnan(g) & np.isnan(r)) | (np.isinf(r) & (g == r))
        finite = np.isfinite(r)
        bad_special = (~finite) & (~same_special)
        with np.errstate(all="ignore"):
            err = np.abs(g - r) / (tol[:, m] + 1e-300)
        err = np.where(finite, err, np.where(bad_special, np.inf, 0.0))
        err = np.where(finite & ~np.isfinite(g), np.inf, err)
        res[name] = float(err.max()) if len(err) else 0.0
    return res


def assert_table_close(got, ref, what=""):
    errs = column_errors(got, ref)
    record(what or "table", errs)
    bad = {k: v for k, v in errs.items() if not v <= 1.0}
    assert not bad, f"{what}: columns outside the parity rule (error / tolerance): {bad}"
    return errs


def bundle_tolerances(ref_rows):
    """Tolerance of the 13 bundle means given the reference rows they average: the mean of the row tolerances
    (for the eigen ratios that is the conditioned rule, row by row)."""
    ref_rows = np.asarray(ref_rows, dtype=np.float64).reshape(-1, len(COLUMNS))
    tol = column_tolerances(ref_rows)
    out = np.empty(len(BUNDLE_SOURCE))
    for j, src in enumerate(BUNDLE_SOURCE):
        m = COLUMNS.index(src)
        col, t = ref_rows[:, m], tol[:, m]
        ok = np.isfinite(col)
        # mean of the finite rows' tolerances, but never below the rule applied to the mean itself
        with np.errstate(all="ignore"):
            floor = RTOL * abs(float(np.mean(col[ok]))) + ATOL[src] if ok.any() else ATOL[src]
            out[j] = max(float(np.mean(t[ok])) if ok.any() else 0.0, floor)
    return out


def assert_bundle_close(got14, ref14, ref_rows=None, what=""):
    """Bundle summary: count bit-exact, 13 means under the same rule (the row tolerances averaged when the
    reference rows are given, else the rule applied to the mean)."""
    got14 = np.asarray(got14, dtype=np.float64)
    ref14 = np.asarray(ref14, dtype=np.float64)
    assert got14[0] == ref14[0], f"{what}: n_streamlines {got14[0]} != {ref14[0]}"
    tols = bundle_tolerances(ref_rows) if ref_rows is not None and len(ref_rows) else None
    errs = {}
    for j, src in enumerate(BUNDLE_SOURCE):
        g, r = got14[1 + j], ref14[1 + j]
        if np.isnan(r) or np.isinf(r):
            assert (np.isnan(g) and np.isnan(r)) or g == r, f"{what}: {src} mean {g} vs {r}"
            continue
        tol = tols[j] if tols is not None else RTOL * abs(r) + ATOL[src]
        errs[src + "_mean"] = abs(g - r) / (tol + 1e-300)
        assert abs(g - r) <= tol, f"{what}: {src} mean {g!r} vs {r!r} (tolerance {tol:.3g})"
    record((what or "bundle") + " [bundle means]", errs)
