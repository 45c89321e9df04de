"""CPU suite, part 2: host-side logic — VTK reader/writer, CSR construction, the drop-in shims,
and that libtractgeom.so loads and exports every symbol include/tractgeom.h declares."""
import ctypes
import gzip
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from lesion_condition_vae_b200 import _lib, synth, vtk_io
from lesion_condition_vae_b200 import tract_geom_proc as tgp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def small():
    rng = np.random.default_rng(5)
    n = np.array([0, 1, 2, 3, 7, 50, 4, 3], dtype=np.int64)
    return synth.random_walk_csr(n, 5)


@pytest.mark.parametrize("binary", [True, False])
@pytest.mark.parametrize("layout", ["classic", "offsets"])
@pytest.mark.parametrize("ptype", ["double", "float"])
def test_vtk_roundtrip(tmp_path, small, binary, layout, ptype):
    pts, off = small
    p = vtk_io.write_polylines(tmp_path / "a.vtk", pts, off, binary=binary, point_dtype=ptype, layout=layout)
    q, o = vtk_io.read_polylines_csr(p)
    assert np.array_equal(o, off)
    assert q.dtype == (np.float64 if ptype == "double" else np.float32)
    assert np.array_equal(q, pts.astype(q.dtype))


def test_vtk_gz_and_connectivity(tmp_path, small):
    pts, off = small
    perm = np.random.default_rng(1).permutation(len(pts))
    shuffled = np.empty_like(pts); shuffled[perm] = pts          # point i stored at perm[i]
    p = vtk_io.write_polylines(tmp_path / "b.vtk.gz", shuffled, off, connectivity=perm)
    with open(p, "rb") as f:
        assert f.read(2) == b"\x1f\x8b"
    q, o = vtk_io.read_polylines_csr(p)
    assert np.array_equal(o, off) and np.array_equal(q, pts)     # gather by index, ref:19-20


def test_vtk_skips_other_sections(tmp_path):
    txt = (b"# vtk DataFile Version 4.2\nx\nASCII\nDATASET POLYDATA\nPOINTS 4 float\n0 0 0 1 0 0\n1 1 0 2 1 1\n"
           b"METADATA\nINFORMATION 1\nNAME L2_NORM_RANGE LOCATION vtkDataArray\nDATA 2 0 2.4\n\n"
           b"VERTICES 1 2\n1 0\nLINES 1 5\n4 0 1 2 3\nPOINT_DATA 4\nSCALARS s float\nLOOKUP_TABLE default\n1 2 3 4\n")
    p = tmp_path / "c.vtk"; p.write_bytes(txt)
    q, o = vtk_io.read_polylines_csr(p)
    assert o.tolist() == [0, 4] and q.shape == (4, 3) and q[3].tolist() == [2, 1, 1]


def test_vtk_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        vtk_io.read_polylines_csr(tmp_path / "missing.vtk")
    p = tmp_path / "bad.vtk"; p.write_bytes(b"hello\n")
    with pytest.raises(vtk_io.VTKFormatError):
        vtk_io.read_polylines_csr(p)


def test_legacy_lines_walk_matches_reference_loop():
    lines = np.array([3, 5, 6, 7, 0, 2, 1, 0, 4, 9, 8, 7, 6], dtype=np.int64)
    off, conn = vtk_io.legacy_lines_to_csr(lines)
    # the reference's while-loop (tract_geom_proc.py:17-25), restated on the index level
    i, exp = 0, []
    while i < len(lines):
        k = int(lines[i]); exp.append(lines[i + 1:i + 1 + k].tolist()); i += 1 + k
    got = [conn[off[s]:off[s + 1]].tolist() for s in range(len(off) - 1)]
    assert got == exp
    off2, conn2 = vtk_io.legacy_lines_to_csr(np.array([2, 0, 1, 2, 2, 3, 2, 4, 5]), 3)   # uniform fast path
    assert off2.tolist() == [0, 2, 4, 6] and conn2.tolist() == [0, 1, 2, 3, 4, 5]


def test_native_ingest_helpers_agree_with_the_numpy_statements(tmp_path, monkeypatch):
    """vtk_io uses the library's host-side helpers when it is built and numpy otherwise: same arrays either way,
    same errors on corrupt input (ASCII numbers incl. signs, exponents, inf / nan; ragged and empty cells)."""
    assert vtk_io._native() is not None                          # the library is built in this checkout
    rng = np.random.default_rng(7)
    lines = []
    for _ in range(300):
        k = int(rng.integers(0, 9))
        lines += [k] + rng.integers(0, 1000, k).tolist()
    lines = np.asarray(lines, dtype=np.int64)
    vals = np.concatenate([rng.normal(size=500) * 10.0 ** rng.integers(-30, 30, 500), [0.0, -0.0, np.inf, -np.inf, 1e-320, 1.7976931348623157e308]])
    txt = " ".join(repr(float(v)) for v in vals).replace("e+", "E+").encode() + b" nan +2.5 -3 7\n 8"
    pts, off = synth.config1(S=40, seed=2)
    files = [vtk_io.write_polylines(tmp_path / f"f{i}.vtk", pts, off, binary=b, layout=l)
             for i, (b, l) in enumerate([(False, "classic"), (False, "offsets"), (True, "classic")])]

    def run():
        f, used = vtk_io._Cursor(txt).ascii(">f8", len(vals) + 4), None
        cur = vtk_io._Cursor(txt); cur.ascii(">f8", len(vals) + 2); ints = cur.ascii(">i4", 3)
        return vtk_io.legacy_lines_to_csr(lines), f, ints, [vtk_io.read_polylines_csr(p) for p in files]

    native = run()
    monkeypatch.setattr(vtk_io, "_NATIVE", None)
    plain = run()
    assert np.array_equal(native[0][0], plain[0][0]) and np.array_equal(native[0][1], plain[0][1])
    assert np.array_equal(native[1], plain[1], equal_nan=True) and np.array_equal(native[1][:len(vals)], vals)
    assert native[2].tolist() == plain[2].tolist() == [-3, 7, 8]
    assert native[1][len(vals) + 1:].tolist() == [2.5, -3.0, 7.0]
    for (a, b), (c, d) in zip(native[3], plain[3]):
        assert np.array_equal(a, c) and np.array_equal(b, d)
    for mode in (False, None):                                   # native, then numpy
        monkeypatch.setattr(vtk_io, "_NATIVE", mode)
        with pytest.raises(vtk_io.VTKFormatError):
            vtk_io.legacy_lines_to_csr(np.array([3, 0, 1], dtype=np.int64))          # count overruns the array
        with pytest.raises(vtk_io.VTKFormatError):
            vtk_io._Cursor(b"1 2 3").ascii(">f8", 4)                                  # truncated


def test_prefix_rule():
    n = np.array([5, 2, 9, 0, 3, 3, 1, 4])
    assert tgp._prefix_for(n, 1) == 1
    assert tgp._prefix_for(n, 2) == 3
    assert tgp._prefix_for(n, 4) == 6
    assert tgp._prefix_for(n, 99) == 8
    assert tgp._prefix_for(n, 1, start=3) == 5


def test_read_streamlines_from_vtk_compat(tmp_path):
    pts, off = synth.lines_to_csr(synth.adversarial_lines())
    p = vtk_io.write_polylines(tmp_path / "adv.vtk", pts, off)
    sls = tgp.read_streamlines_from_vtk(p)
    assert len(sls) == 11                       # n=2, NaN and inf lines dropped by the loader filter
    assert len(tgp.read_streamlines_from_vtk(p, max_streamlines=4)) == 4


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "tractgeom.h")).read()
    declared = set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.tg_abi_version() == 3
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", nm), name


def test_library_is_built_from_these_sources():
    """the binary the tests and the bench map must be the one HEAD's sources produce.  The id
    compiled into the library (file bytes AND the loaded image) equals the sha256 id of csrc/* + tractgeom.h +
    the nvcc command line; editing any kernel header changes the id (so a stale binary cannot go unnoticed)."""
    from lesion_condition_vae_b200 import build as _b
    if os.environ.get("TG_LIB"):
        pytest.skip("tuning variant selected with TG_LIB")
    assert _b.library_id() == _b.source_id() == _lib.build_id()
    deps = {os.path.basename(p) for p in _b.dependency_files()}
    assert {"tg_kernels.cu", "tg_grouped.cuh", "tg_device.cuh", "tractgeom.h"} <= deps
    # every #include "..." of the translation unit is a tracked dependency
    for path in _b.dependency_files():
        for inc in re.findall(r'#include\s+"([^"]+)"', open(path).read()):
            assert os.path.basename(inc) in deps, (path, inc)
    # no stray library variants in the package (only the one the id check covers ships)
    pkg = os.path.dirname(_b.LIB)
    assert [f for f in os.listdir(pkg) if f.endswith(".so")] == ["libtractgeom.so"]


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.TractGeomError) as e:
        _lib.Context(0)
    assert e.value.code == -4
    pts, off = synth.config1(S=5)
    with pytest.raises(_lib.TractGeomError):
        tgp.compute_streamline_metrics_csr(pts, off)


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "lesion_condition_vae_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "reference_runner" not in src and "streamline_oracle" not in src and "resample_oracle" not in src, f


def test_new_entry_points_reject_a_null_context_without_touching_cuda():
    """ABI v2 additions validate their arguments before any CUDA call (runs on a box with no GPU)."""
    lib = _lib.load()
    assert lib.tg_bundle_spread_dev(None, None, None, None, 0, None, 0, None, None, None, None) == -1
    assert lib.tg_metrics_csr_host_ex(None, None, 0, None, 0, 0, None, 0, None, None, None, None, None) == -1
    assert lib.tg_resample_csr_dev(None, None, 0, None, 0, 0, 100, None, None) == -1
    assert lib.tg_resample_csr_host(None, None, 0, None, 0, 0, 100, None) == -1
    assert b"null context" in lib.tg_last_error()


# ---- reader hardening: byte-level fixtures written WITHOUT this package's writer ------------------------
VTK_FIXTURES = os.path.join(ROOT, "tests", "golden", "vtk")


@pytest.mark.parametrize("name,case", [("v42_field_float.vtk", "v42_field_float"), ("v51_offsets.vtk", "v51_offsets"),
                                       ("v51_offsets.vtk.gz", "v51_offsets"), ("v30_ascii_field.vtk", "v30_ascii_field")])
def test_reader_on_vtkpolydatawriter_layouts(name, case):
    """FIELD data in front of the geometry (v4.2 binary float, and ASCII), v5.1 OFFSETS / CONNECTIVITY vtktypeint64 with a
    METADATA block, the same gzipped, trailing CELL_DATA / POINT_DATA, and a connectivity that is not the identity."""
    exp = np.load(os.path.join(VTK_FIXTURES, "expected.npz"))
    pts, off = vtk_io.read_polylines_csr(os.path.join(VTK_FIXTURES, name), dtype=np.float64)
    assert np.array_equal(off, exp[case + "/offsets"]) and np.array_equal(pts, exp[case + "/points"])
    # the device-path reader: same values, binary files left in the file's big-endian storage (no per-point host work)
    raw, off2 = vtk_io.read_polylines_raw(os.path.join(VTK_FIXTURES, name))
    assert np.array_equal(off2, off) and np.array_equal(raw.astype(np.float64), pts)
    if not name.startswith("v30"):
        assert raw.dtype.byteorder == ">" and _lib.dtype_code(raw) in (_lib.F32_BE, _lib.F64_BE)


def test_raw_reader_streams_the_points_block_into_a_given_buffer(tmp_path):
    """read_polylines_raw(arena=...) reads the POINTS block of an uncompressed binary file straight into the arena (readinto),
    larger than the 64 KB header probe, and copies a gunzipped block into it."""
    class Arena:                                           # stands in for _lib.PinnedArena (which needs a CUDA device)
        def __init__(self):
            self.blocks = []

        def take(self, nbytes, dtype=np.uint8):
            self.blocks.append(np.zeros(nbytes, dtype=np.uint8))
            return self.blocks[-1].view(dtype)

    pts, off = synth.random_walk_csr(synth.lengths_uniform(np.random.default_rng(3), 400, 3, 90), 8)
    for gz, pd_ in ((False, "float"), (False, "double"), (True, "float")):
        p = vtk_io.write_polylines(tmp_path / ("a.vtk.gz" if gz else f"a_{pd_}.vtk"), pts, off, binary=True, point_dtype=pd_)
        ar = Arena()
        raw, o = vtk_io.read_polylines_raw(p, ar)
        ref, o_ref = vtk_io.read_polylines_csr(p)
        assert np.array_equal(o, o_ref) and np.array_equal(raw.astype(ref.dtype), ref)
        assert len(ar.blocks) == 1 and np.shares_memory(raw, ar.blocks[0]) and ar.blocks[0].nbytes == ref.nbytes > (1 << 16)


def test_corrupt_cell_array_is_rejected_not_looped_on(monkeypatch):
    """a negative cell count never advanced the numpy fallback walk.  Both walks (native helper, numpy) reject it."""
    bad = np.array([3, 0, 1, 2, -1, 5, 6], dtype=np.int64)
    with pytest.raises(vtk_io.VTKFormatError):
        vtk_io.legacy_lines_to_csr(bad)
    monkeypatch.setattr(vtk_io, "_NATIVE", None)
    with pytest.raises(vtk_io.VTKFormatError):
        vtk_io.legacy_lines_to_csr(bad)
    with pytest.raises(vtk_io.VTKFormatError):
        vtk_io.legacy_lines_to_csr(np.array([2, 0, 1, 4, 2], dtype=np.int64))       # last count overruns
    off, conn = vtk_io.legacy_lines_to_csr(np.array([2, 0, 1, 3, 2, 3, 4], dtype=np.int64))
    assert off.tolist() == [0, 2, 5] and conn.tolist() == [0, 1, 2, 3, 4]


def test_select_prefix_matches_the_reference_loader_rule():
    """The lazy prefix scan == the loader loop of tract_geom_proc.py:17-25, on native and on big-endian storage."""
    from lesion_condition_vae_b200 import tract_driver as td
    rng = np.random.default_rng(12)
    for trial in range(30):
        n = rng.integers(0, 7, size=rng.integers(1, 60))
        pts, off = synth.random_walk_csr(np.maximum(n, 1), 100 + trial)
        off = np.concatenate([[0], np.cumsum(np.maximum(n, 1))]).astype(np.int64)
        for s in rng.integers(0, len(n), size=3):
            if rng.random() < 0.5 and off[s + 1] > off[s]:
                pts[off[s] + rng.integers(0, off[s + 1] - off[s]), rng.integers(0, 3)] = [np.nan, np.inf][trial % 2]
        for ms in (None, 1, 3, 10, 0):
            want = []
            for s in range(len(n)):
                sl = pts[off[s]:off[s + 1]]
                if sl.shape[0] > 2 and np.isfinite(sl).all():
                    want.append(s)
                    if ms is not None and len(want) >= ms:
                        break
            for view in (pts, pts.astype(">f8"), pts.astype(">f4")):
                ref = want if view.dtype.itemsize == 8 else None
                got = td.select_prefix(view, off, ms).tolist()
                if ref is not None:
                    assert got == ref, (trial, ms)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) emits one JSON line with the contract keys,
    the same metric / unit as the GPU arm, a cpu_baseline describing the run and a zero-copy e2e record."""
    import json
    import sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--config", "0"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "streamlines_per_sec" and line["unit"] == "streamlines/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] and "workload" in line["config"]
