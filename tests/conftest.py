import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The C-ABI library is git-ignored.  It is (re)built whenever the build id in the file differs from the id
    of the sources on disk (build.source_id: sha256 over csrc/*, include/tractgeom.h and the nvcc command line),
    so the binary under test is the one these sources produce.  On the GPU box the snapshot's prebuilt .so has
    the matching id and is used as is; a mismatch there with no nvcc is an error, not a skip."""
    from lesion_condition_vae_b200 import build as _b
    if os.environ.get("TG_LIB"):                               # an explicitly selected tuning variant
        return
    if _b.is_stale():
        try:
            _b.find_nvcc()
        except RuntimeError:
            if _b.library_id() is None:
                print("[conftest] no library and no nvcc: tests that need the library will fail")
                return
            raise RuntimeError(f"libtractgeom.so has build id {_b.library_id()} but the sources are {_b.source_id()} and nvcc is missing")
        _b.build()
    from lesion_condition_vae_b200 import _lib
    assert _lib.build_id() == _b.source_id(), f"loaded library {_lib.build_id()} != sources {_b.source_id()}"
    print(f"[conftest] libtractgeom.so build id {_lib.build_id()}")


def pytest_sessionfinish(session, exitstatus):
    """Dump the parity budget (worst error / tolerance per column and case, collected by parity_rules.record) of a
    GPU session to gpurun_out/parity_budget.json; tools/parity_budget.py turns it into profiles/parity_budget_r2.json."""
    try:
        import json
        import parity_rules
        if not parity_rules.BUDGET or not getattr(session.config.option, "markexpr", "") == "gpu":
            return
        from lesion_condition_vae_b200 import _lib
        out = os.path.join(ROOT, "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_budget.json"), "w") as f:
            json.dump({"build_id": _lib.build_id(), "exitstatus": int(exitstatus), "rule": "error / tolerance, <= 1 passes (tests/parity_rules.py)",
                       "cases": parity_rules.BUDGET}, f, indent=1, sort_keys=True)
    except Exception as e:                                       # never turn a reporting problem into a test failure
        print(f"[conftest] parity budget not written: {e}")


@pytest.fixture(autouse=True)
def release_torch_cache(request):
    """The library allocates with cudaMalloc, outside torch's caching allocator: hand the cache back after every
    GPU test so a 100 GB full-size test cannot starve the next one."""
    yield
    if "gpu" in request.keywords:
        import torch
        if torch.cuda.is_available():
            torch.cuda.empty_cache()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    z = np.load(path)
    cases = {}
    for key in z.files:
        name, field = key.split("/", 1)
        cases.setdefault(name, {})[field] = z[key]
    for name, c in cases.items():
        if "input_of" in c:
            src = cases[str(c["input_of"])]
            c["points"], c["offsets"] = src["points"], src["offsets"]
        ms = int(c["max_streamlines"])
        c["max_streamlines"] = None if ms < 0 else ms
    return cases


@pytest.fixture(scope="session")
def gpu_ctx():
    """Context on cuda:0.  Fails (does not skip) when the library or device is missing: a GPU test
    that silently passes without the native code would be worthless."""
    from lesion_condition_vae_b200 import _lib
    return _lib.default_context(0)
