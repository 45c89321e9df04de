"""The parity rule shared by every comparison against the oracle.

    |got - ref| <= RTOL * |ref| + ATOL[metric]        inf must equal inf, NaN must equal NaN

RTOL = 1e-9 is BASELINE.json's contract.  ATOL exists only because several reference outputs are
rounding noise on degenerate inputs (straight line: curv_mean 1e-16, ang_dispersion 2e-31, ...)
or sums that cancel (torsion_mean, centroid near 0), where a relative rule is meaningless.
For the two eigenvalue ratios the oracle's own LAPACK error is ~1e-16 * lambda1/lambda3
(SURVEY.md F6), so their RTOL is widened to 2e-14 * (lambda1/lambda3) once that exceeds 1e-9.
"""
import numpy as np

RTOL = 1e-9
COLUMNS = (
    "length", "end_to_end", "tortuosity", "straightness", "curv_mean", "curv_std", "curv_energy",
    "torsion_mean", "bend_angle_mean", "bbox_vol", "elongation_ratio", "planarity_ratio",
    "anisotropy_ratio", "centroid_x", "centroid_y", "centroid_z", "ang_dispersion",
)
ATOL = {
    "length": 0.0, "end_to_end": 1e-15, "tortuosity": 0.0, "straightness": 0.0,
    "curv_mean": 1e-13, "curv_std": 1e-13, "curv_energy": 1e-13, "torsion_mean": 1e-12,
    "bend_angle_mean": 2e-9, "bbox_vol": 1e-12, "elongation_ratio": 0.0, "planarity_ratio": 0.0,
    "anisotropy_ratio": 0.0, "centroid_x": 1e-12, "centroid_y": 1e-12, "centroid_z": 1e-12,
    "ang_dispersion": 1e-14,
}
BUNDLE_SOURCE = ("length", "tortuosity", "curv_mean", "curv_energy", "torsion_mean", "bend_angle_mean",
                 "elongation_ratio", "planarity_ratio", "anisotropy_ratio", "ang_dispersion",
                 "centroid_x", "centroid_y", "centroid_z")


def column_errors(got, ref):
    """Per-column max of |got-ref| / (RTOL*|ref| + ATOL) over rows (<= 1 passes); got/ref (R,17)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    res = {}
    with np.errstate(all="ignore"):
        cond = ref[:, 10] * ref[:, 11]          # lambda1/lambda3 = elongation * planarity
    for m, name in enumerate(COLUMNS):
        g, r = got[:, m], ref[:, m]
        same_special = (np.isnan(g) & np.isnan(r)) | (np.isinf(r) & (g == r))
        finite = np.isfinite(r)
        bad_special = (~finite) & (~same_special)
        rtol = np.full(len(r), RTOL)
        if name in ("elongation_ratio", "planarity_ratio"):
            with np.errstate(all="ignore"):
                rtol = np.where(np.isfinite(cond), np.maximum(RTOL, 2e-14 * cond), RTOL)
        with np.errstate(all="ignore"):
            err = np.abs(g - r) / (rtol * np.abs(r) + ATOL[name] + 1e-300)
        err = np.where(finite, err, np.where(bad_special, np.inf, 0.0))
        err = np.where(finite & ~np.isfinite(g), np.inf, err)
        res[name] = float(err.max()) if len(err) else 0.0
    return res


def assert_table_close(got, ref, what=""):
    errs = column_errors(got, ref)
    bad = {k: v for k, v in errs.items() if not v <= 1.0}
    assert not bad, f"{what}: columns outside the 1e-9 parity rule (error / tolerance): {bad}"
    return errs


def assert_bundle_close(got14, ref14, ref_rows=None, what=""):
    """Bundle summary: count bit-exact, 13 means under the same rule (ATOL of the source column)."""
    got14 = np.asarray(got14, dtype=np.float64)
    ref14 = np.asarray(ref14, dtype=np.float64)
    assert got14[0] == ref14[0], f"{what}: n_streamlines {got14[0]} != {ref14[0]}"
    for j, src in enumerate(BUNDLE_SOURCE):
        g, r = got14[1 + j], ref14[1 + j]
        if np.isnan(r) or np.isinf(r):
            assert (np.isnan(g) and np.isnan(r)) or g == r, f"{what}: {src} mean {g} vs {r}"
            continue
        rtol = RTOL
        if src in ("elongation_ratio", "planarity_ratio") and ref_rows is not None and len(ref_rows):
            with np.errstate(all="ignore"):
                cond = ref_rows[:, 10] * ref_rows[:, 11]
            cond = cond[np.isfinite(cond)]
            if len(cond):
                rtol = max(RTOL, 2e-14 * float(cond.max()))
        assert abs(g - r) <= rtol * abs(r) + ATOL[src], f"{what}: {src} mean {g!r} vs {r!r}"
