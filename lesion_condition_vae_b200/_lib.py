"""ctypes binding of include/tractgeom.h.  No fallback: a missing library or device raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TG_LIB") or os.path.join(HERE, "libtractgeom.so")   # TG_LIB: a tuning variant built by build.py

N_METRICS = 17
N_BUNDLE_COLS = 13
KEEP_LOADER, KEEP_LENGTH, KEEP_BOTH = 1, 2, 3
F64, F32, F64_BE, F32_BE = 0, 1, 2, 3          # tg_dtype: native float64 / float32, big-endian float64 / float32 (binary VTK)

# every symbol include/tractgeom.h declares (tests check the .so exports exactly these)
EXPORTS = (
    "tg_abi_version", "tg_build_id", "tg_last_error", "tg_device_count", "tg_create", "tg_destroy", "tg_synchronize",
    "tg_stream", "tg_host_alloc", "tg_host_free", "tg_metrics_csr_dev", "tg_bundle_reduce_dev",
    "tg_metrics_csr_host", "tg_launch_count", "tg_bundle_spread_dev", "tg_metrics_csr_host_ex",
    "tg_resample_csr_dev", "tg_resample_csr_host", "tg_bundle_partials_dev",
    "tg_batch_begin", "tg_batch_push", "tg_batch_run", "tg_batch_size",
    "tg_vtk_lines_to_csr", "tg_vtk_cells_be32_to_csr", "tg_parse_ascii_f64", "tg_parse_ascii_i64",
)


class TractGeomError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tractgeom error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """dlopen libtractgeom.so and declare the prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m lesion_condition_vae_b200.build` "
            "(nvcc, sm_100a).  There is no CPU implementation to fall back to.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.tg_abi_version.restype = i32
    lib.tg_last_error.restype = C.c_char_p
    lib.tg_build_id.restype = C.c_char_p
    lib.tg_device_count.argtypes = [C.POINTER(i32)]
    lib.tg_create.argtypes = [i32, C.POINTER(vp)]
    lib.tg_destroy.argtypes = [vp]
    lib.tg_synchronize.argtypes = [vp]
    lib.tg_stream.argtypes = [vp, C.POINTER(vp)]
    lib.tg_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.tg_host_free.argtypes = [vp]
    lib.tg_metrics_csr_dev.argtypes = [vp, vp, i32, vp, i64, i64, vp, vp, vp]
    lib.tg_bundle_reduce_dev.argtypes = [vp, vp, vp, vp, i64, vp, i64, vp, vp, vp]
    lib.tg_batch_begin.argtypes = [vp, i64, i64]
    lib.tg_batch_push.argtypes = [vp, vp, i32, i64, vp, i64]
    lib.tg_batch_run.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp]
    lib.tg_batch_size.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.tg_bundle_partials_dev.argtypes = [vp, vp, vp, vp, i64, vp, i64, vp, vp]
    lib.tg_metrics_csr_host.argtypes = [vp, vp, i32, vp, i64, i64, vp, i64, vp, vp, vp, vp]
    lib.tg_launch_count.argtypes = [vp, C.POINTER(i64)]
    lib.tg_bundle_spread_dev.argtypes = [vp, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp]
    lib.tg_resample_csr_dev.argtypes = [vp, vp, i32, vp, i64, i64, i32, vp, vp]
    lib.tg_resample_csr_host.argtypes = [vp, vp, i32, vp, i64, i64, i32, vp]
    lib.tg_vtk_lines_to_csr.argtypes = [vp, i64, vp, vp, C.POINTER(i64), C.POINTER(i64)]
    lib.tg_vtk_cells_be32_to_csr.argtypes = [vp, i64, vp, vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i32)]
    lib.tg_parse_ascii_f64.argtypes = [C.c_char_p, i64, i64, vp, C.POINTER(i64)]
    lib.tg_parse_ascii_i64.argtypes = [C.c_char_p, i64, i64, vp, C.POINTER(i64)]
    lib.tg_metrics_csr_host_ex.argtypes = [vp, vp, i32, vp, i64, i64, vp, i64, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if name not in ("tg_abi_version", "tg_last_error", "tg_build_id"):
            fn.restype = i32
    _lib = lib
    return lib


def build_id():
    """Build id compiled into the LOADED library (build.source_id() of the sources it was made from)."""
    return load().tg_build_id().decode()


def check(rc):
    if rc != 0:
        raise TractGeomError(rc, load().tg_last_error().decode(errors="replace"))


def dtype_code(a):
    """tg_dtype of a numpy point array, or None when it has to be converted to native float64 first."""
    dt = a.dtype
    if dt.kind != "f" or dt.itemsize not in (4, 8):
        return None
    big = dt.byteorder == ">" or (dt.byteorder == "=" and not np.little_endian)
    return {(8, False): F64, (4, False): F32, (8, True): F64_BE, (4, True): F32_BE}[(dt.itemsize, big)]


def _ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One tg_context: a CUDA device, a stream and grow-only scratch.  Not thread-safe."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        self._lib = load()
        check(self._lib.tg_create(int(device), C.byref(self._h)))
        self.device = int(device)

    def close(self):
        if self._h:
            self._lib.tg_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._lib.tg_synchronize(self._h))

    @property
    def stream(self):
        s = C.c_void_p()
        check(self._lib.tg_stream(self._h, C.byref(s)))
        return s.value or 0

    @property
    def launches(self):
        n = C.c_int64()
        check(self._lib.tg_launch_count(self._h, C.byref(n)))
        return n.value

    # ---- device-pointer calls (ints are raw device addresses, e.g. torch.Tensor.data_ptr()) ----
    def metrics_dev(self, d_xyz, dtype_code, d_offsets, S, P, d_out, d_keep, stream=0):
        check(self._lib.tg_metrics_csr_dev(self._h, d_xyz, dtype_code, d_offsets, S, P, d_out, d_keep, stream or None))

    def bundle_reduce_dev(self, d_out, d_keep, d_select, S, bundle_offsets, d_sums, d_counts, stream=0):
        bo = np.ascontiguousarray(bundle_offsets, dtype=np.int64)
        check(self._lib.tg_bundle_reduce_dev(self._h, d_out, d_keep, d_select or None, S, _ptr(bo), len(bo) - 1,
                                             d_sums, d_counts, stream or None))

    def bundle_partials_dev(self, d_out, d_keep, d_select, S, bundle_offsets, d_partials, stream=0):
        """Bundle partial moments as (B,27) float64 rows {13 sums | kept rows | 13 non-NaN counts}: the all-gather payload."""
        bo = np.ascontiguousarray(bundle_offsets, dtype=np.int64)
        check(self._lib.tg_bundle_partials_dev(self._h, d_out, d_keep, d_select or None, S, _ptr(bo), len(bo) - 1,
                                               d_partials, stream or None))

    def bundle_spread_dev(self, d_out, d_keep, d_select, S, bundle_offsets, d_sums, d_counts, d_spread, stream=0):
        """Opt-in {std, min, max} of the 13 bundle columns; call after :meth:`bundle_reduce_dev` (same arguments)."""
        bo = np.ascontiguousarray(bundle_offsets, dtype=np.int64)
        check(self._lib.tg_bundle_spread_dev(self._h, d_out, d_keep, d_select or None, S, _ptr(bo), len(bo) - 1,
                                             d_sums, d_counts, d_spread, stream or None))

    def resample_dev(self, d_xyz, dtype_code, d_offsets, S, P, n_nodes, d_nodes, stream=0):
        check(self._lib.tg_resample_csr_dev(self._h, d_xyz, dtype_code, d_offsets, S, P, n_nodes, d_nodes, stream or None))

    def resample_host(self, points, offsets, n_nodes=100, nodes=None):
        """Arc-length resampling to ``n_nodes`` points per polyline -> float64 (S, n_nodes, 3)."""
        points = np.ascontiguousarray(points)
        if points.dtype == np.float32:
            code = F32
        else:
            points = np.ascontiguousarray(points, dtype=np.float64)
            code = F64
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        S = len(offsets) - 1
        P = points.shape[0] if points.ndim == 2 else points.size // 3
        if nodes is None:
            nodes = np.empty((S, n_nodes, 3), dtype=np.float64)
        assert nodes.shape == (S, n_nodes, 3) and nodes.dtype == np.float64 and nodes.flags.c_contiguous
        check(self._lib.tg_resample_csr_host(self._h, _ptr(points), code, _ptr(offsets), S, P, int(n_nodes), _ptr(nodes)))
        return nodes

    # ---- host-buffer call: H2D + kernels + D2H, synchronous ----
    def metrics_host(self, points, offsets, bundle_offsets=None, want_rows=True, out=None, keep=None, spread=None):
        """points (P,3) float64|float32 C-contiguous, offsets int64[S+1].

        ``out`` / ``keep``: optional preallocated result arrays (e.g. pinned memory).
        ``spread``: optional float64 (B,13,3) array that receives the opt-in {std, min, max} per bundle column.
        Returns (out (17,S) float64 or None, keep uint8[S], sums (B,13), counts (B,14))."""
        points = np.ascontiguousarray(points)
        code = dtype_code(points)                       # big-endian float32 / float64 (binary VTK) go to the device as they are
        if code is None:
            points = points.astype(np.float64)
            code = F64
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        S = len(offsets) - 1
        P = points.shape[0] if points.ndim == 2 else points.size // 3
        if bundle_offsets is None:
            bundle_offsets = np.array([0, S], dtype=np.int64)
        bo = np.ascontiguousarray(bundle_offsets, dtype=np.int64)
        B = len(bo) - 1
        if out is None:
            out = np.empty((N_METRICS, S), dtype=np.float64) if want_rows else None
        if keep is None:
            keep = np.empty(S, dtype=np.uint8)
        assert out is None or (out.shape == (N_METRICS, S) and out.dtype == np.float64)
        assert keep.shape == (S,) and keep.dtype == np.uint8
        sums = np.empty((B, N_BUNDLE_COLS), dtype=np.float64)
        counts = np.empty((B, N_BUNDLE_COLS + 1), dtype=np.int64)
        if spread is not None:
            assert spread.shape == (B, N_BUNDLE_COLS, 3) and spread.dtype == np.float64 and spread.flags.c_contiguous
            check(self._lib.tg_metrics_csr_host_ex(self._h, _ptr(points), code, _ptr(offsets), S, P, _ptr(bo), B,
                                                   _ptr(out), _ptr(keep), _ptr(sums), _ptr(counts), _ptr(spread)))
            return out, keep, sums, counts
        check(self._lib.tg_metrics_csr_host(self._h, _ptr(points), code, _ptr(offsets), S, P, _ptr(bo), B,
                                            _ptr(out), _ptr(keep), _ptr(sums), _ptr(counts)))
        return out, keep, sums, counts


    # ---- batch of files: push as parsed (H2D overlaps the parsing of the next file), run once ----
    def batch_begin(self, P_capacity, S_capacity):
        check(self._lib.tg_batch_begin(self._h, int(P_capacity), int(S_capacity)))

    def batch_push(self, points, offsets):
        """One file's polylines: points (P,3) in any supported storage (see dtype_code), local offsets starting at 0.
        The array must stay alive and unchanged until batch_run returns (pinned memory makes the copy asynchronous)."""
        code = dtype_code(points)
        if code is None or not points.flags.c_contiguous:
            points = np.ascontiguousarray(points, dtype=np.float64)
            code = F64
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        S = len(offsets) - 1
        P = points.shape[0] if points.ndim == 2 else points.size // 3
        check(self._lib.tg_batch_push(self._h, _ptr(points), code, P, _ptr(offsets), S))
        return points

    def batch_run(self, bundle_offsets, want_rows=False, spread=None):
        """-> (out (17,S) or None, keep uint8[S], sums (B,13), counts (B,14)) over everything pushed since batch_begin."""
        S, P = C.c_int64(), C.c_int64()
        check(self._lib.tg_batch_size(self._h, C.byref(S), C.byref(P)))
        S = S.value
        bo = np.ascontiguousarray(bundle_offsets, dtype=np.int64)
        B = len(bo) - 1
        out = np.empty((N_METRICS, S), dtype=np.float64) if want_rows else None
        keep = np.empty(S, dtype=np.uint8)
        sums = np.empty((B, N_BUNDLE_COLS), dtype=np.float64)
        counts = np.empty((B, N_BUNDLE_COLS + 1), dtype=np.int64)
        if spread is not None:
            assert spread.shape == (B, N_BUNDLE_COLS, 3) and spread.dtype == np.float64 and spread.flags.c_contiguous
        check(self._lib.tg_batch_run(self._h, _ptr(bo), B, _ptr(out), _ptr(keep), _ptr(sums), _ptr(counts), _ptr(spread)))
        return out, keep, sums, counts


class PinnedArena:
    """Grow-only pinned host memory (tg_host_alloc) handed out as numpy views: the loader parses tract files straight
    into it, so the host-to-device copies of the hot path are real DMA transfers (53 GB/s against ~10 GB/s from
    pageable memory, tools/latency_probe.py) and return at once.  reset() recycles the space; views taken before a
    reset must not be used afterwards."""

    def __init__(self, nbytes=1 << 20):
        import threading
        self._lock = threading.Lock()          # take() is called from the loader's parser threads
        self._lib = load()
        self._blocks = []          # (pointer, capacity, ctypes array); older blocks stay alive until close(): views point into them
        self._cap = 0
        self._used = 0
        self._grow(nbytes)

    def _grow(self, nbytes):
        cap = max(int(nbytes), 2 * self._cap, 1 << 20)
        p = C.c_void_p()
        check(self._lib.tg_host_alloc(C.byref(p), cap))
        self._blocks.append((p, cap, (C.c_ubyte * cap).from_address(p.value)))
        self._cap, self._used = cap, 0

    def reset(self):
        """Recycle.  If the last round needed several blocks, they are replaced by ONE block of their total size, so that
        the same workload fits without another (slow) pinned allocation next time."""
        with self._lock:
            if len(self._blocks) > 1:
                total = sum(cap for _, cap, _ in self._blocks)
                for p, _, _ in self._blocks:
                    self._lib.tg_host_free(p)
                self._blocks, self._cap = [], 0
                self._grow(total)
            self._used = 0

    def take(self, nbytes, dtype=np.uint8):
        """A pinned array of ``nbytes`` bytes viewed as ``dtype`` (64-byte aligned)."""
        with self._lock:
            start = (self._used + 63) & ~63
            if start + nbytes > self._cap:
                self._grow(nbytes + 64)
                start = 0
            self._used = start + int(nbytes)
            raw = np.frombuffer(self._blocks[-1][2], dtype=np.uint8, count=int(nbytes), offset=start)
        return raw.view(dtype)

    def close(self):
        for p, _, _ in self._blocks:
            self._lib.tg_host_free(p)
        self._blocks = []
        self._cap = self._used = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_arena = None


def default_arena():
    """Process-wide pinned arena of the loaders (created on first use; needs a CUDA device)."""
    global _arena
    if _arena is None:
        _arena = PinnedArena()
    return _arena


# ---- host-side ingest helpers (no device needed) ----
def vtk_lines_to_csr(lines):
    """Legacy cell array -> (offsets int64[S+1], connectivity int64[C]); raises TractGeomError on a corrupt array."""
    lib = load()
    lines = np.ascontiguousarray(lines, dtype=np.int64)
    L = lines.size
    offsets = np.empty(L + 1, dtype=np.int64)
    conn = np.empty(max(L, 1), dtype=np.int64)
    ns, nc = C.c_int64(), C.c_int64()
    check(lib.tg_vtk_lines_to_csr(_ptr(lines), L, _ptr(offsets), _ptr(conn), C.byref(ns), C.byref(nc)))
    return offsets[:ns.value + 1].copy(), conn[:nc.value].copy()


def vtk_cells_be32_to_csr(cells):
    """Big-endian int32 classic cell array (as read from a binary file) -> (offsets int64[S+1], connectivity int64[C] or
    None when it is the identity).  One native pass; raises TractGeomError on a corrupt array."""
    lib = load()
    cells = np.ascontiguousarray(cells)
    assert cells.dtype.itemsize == 4
    L = cells.size
    offsets = np.empty(L + 1, dtype=np.int64)
    ns, nc, ident = C.c_int64(), C.c_int64(), C.c_int()
    conn = None
    rc = lib.tg_vtk_cells_be32_to_csr(_ptr(cells), L, _ptr(offsets), None, C.byref(ns), C.byref(nc), C.byref(ident))
    if rc != 0 and b"not the identity" in lib.tg_last_error():
        conn = np.empty(max(L, 1), dtype=np.int64)
        rc = lib.tg_vtk_cells_be32_to_csr(_ptr(cells), L, _ptr(offsets), _ptr(conn), C.byref(ns), C.byref(nc), C.byref(ident))
    check(rc)
    offsets = offsets[:ns.value + 1].copy()
    return offsets, (None if ident.value else conn[:nc.value].copy())


def parse_ascii(buf, start, count, integer=False):
    """`count` whitespace-separated numbers of bytes object `buf` from byte `start` -> (array, bytes consumed)."""
    lib = load()
    out = np.empty(count, dtype=np.int64 if integer else np.float64)
    used = C.c_int64()
    view = memoryview(buf)[start:]
    cbuf = (C.c_char * len(view)).from_buffer_copy(view) if not isinstance(buf, bytes) else None
    fn = lib.tg_parse_ascii_i64 if integer else lib.tg_parse_ascii_f64
    if cbuf is None:
        # bytes: pass a pointer into the object itself (no copy of the remaining file)
        addr = C.cast(C.c_char_p(buf), C.c_void_p).value + start
        check(fn(C.cast(addr, C.c_char_p), len(buf) - start, count, _ptr(out), C.byref(used)))
    else:
        check(fn(cbuf, len(view), count, _ptr(out), C.byref(used)))
    return out, used.value


_default_ctx = {}


def default_context(device=0):
    """Process-wide context per device, created lazily and reused (the reference driver calls the
    hot path 2,368 times per run)."""
    ctx = _default_ctx.get(device)
    if ctx is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx
