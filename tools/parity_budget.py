#!/usr/bin/env python3
"""gpurun_out/parity_budget.json (dumped by a `pytest -m gpu` session, tests/conftest.py) -> profiles/parity_budget_r2.json:
the OBSERVED worst error / tolerance per column, per case and overall, so the margin the parity rules leave is on record.
usage: python tools/parity_budget.py [in.json] [out.json]"""
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_budget.json"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/parity_budget_r2.json"
raw = json.load(open(src))
cases = raw["cases"]
overall = {}
for case, cols in cases.items():
    for c, v in cols.items():
        if c.startswith("max ") or c.startswith("share "):
            continue
        if v > overall.get(c, (0.0, ""))[0]:
            overall[c] = (v, case)
out = {
    "what": "observed max |got - ref| / tolerance (<= 1 passes) of the CUDA path, per df_sl column (and per bundle mean), under tests/parity_rules.py: "
            "1e-9 relative + the per-column noise floor ATOL; eigen ratios per SURVEY.md N7",
    "build_id": raw.get("build_id"),
    "atol": None,
    "worst_per_column": {c: {"error_over_tolerance": round(v, 6), "case": case} for c, (v, case) in sorted(overall.items())},
    "cases": {case: {c: (round(v, 6) if isinstance(v, float) else v) for c, v in sorted(cols.items())} for case, cols in sorted(cases.items())},
}
try:
    sys.path.insert(0, "tests")
    import parity_rules
    out["atol"] = parity_rules.ATOL
    out["rtol"] = parity_rules.RTOL
except Exception:
    pass
json.dump(out, open(dst, "w"), indent=1)
print("worst per column:")
for c, (v, case) in sorted(overall.items(), key=lambda kv: -kv[1][0]):
    print(f"  {c:34s} {v:10.4g}   {case}")
