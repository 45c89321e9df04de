"""Legacy-VTK POLYDATA polylines <-> CSR tractogram, in vectorised numpy.

Stands in for the two steps the reference does before any arithmetic
(/root/reference/src/geometry/tract_geom_proc.py:10-20): ``pv.read`` (points + legacy ``lines``)
and the per-line Python gather ``points[idx]``.  Here the file is parsed once into
``points (P,3)`` + ``connectivity`` + ``offsets`` and gathered in one fancy-index (or not at all
when the connectivity is the identity, the usual case for tractography exports).

Supported: ASCII and BINARY (big-endian) legacy files, ``POINTS n float|double``, the classic
``LINES n size`` cell layout and the VTK 5.1 ``OFFSETS``/``CONNECTIVITY`` layout, ``.vtk.gz``
read straight from memory (the reference driver gunzips to a sibling file first,
comprehensive_tract_geometry_analysis.py:54-76).  Other cell sections (VERTICES, POLYGONS,
TRIANGLE_STRIPS) and METADATA blocks are skipped; attribute data after the geometry is ignored.
If ``pyvista`` is importable it is used for every other format (.vtp, ...).
"""
from __future__ import annotations

import gzip
import os

import numpy as np

_VTK_TYPES = {
    b"float": ">f4", b"double": ">f8", b"int": ">i4", b"unsigned_int": ">u4", b"long": ">i8",
    b"unsigned_long": ">u8", b"vtktypeint64": ">i8", b"vtktypeint32": ">i4", b"short": ">i2",
    b"unsigned_short": ">u2", b"char": ">i1", b"unsigned_char": ">u1", b"vtkidtype": ">i8",
}
_CELL_SECTIONS = (b"VERTICES", b"LINES", b"POLYGONS", b"TRIANGLE_STRIPS")


class VTKFormatError(ValueError):
    pass


_NATIVE = False          # False: not probed yet; None: unavailable


def _native():
    """libtractgeom.so's host-side ingest helpers (tg_vtk_lines_to_csr, tg_parse_ascii_*), or None when the
    library is not built: the numpy statements below then do the same work, slower.  File parsing only — the
    metrics themselves have no CPU path."""
    global _NATIVE
    if _NATIVE is False:
        try:
            from . import _lib
            _lib.load()
            _NATIVE = _lib
        except (ImportError, OSError, AttributeError):
            _NATIVE = None
    return _NATIVE


class _Cursor:
    def __init__(self, buf: bytes):
        self.b = buf
        self.i = 0

    def line(self):
        """Next non-blank line (stripped), or None at EOF."""
        b = self.b
        while self.i < len(b):
            j = b.find(b"\n", self.i)
            if j < 0:
                j = len(b)
            ln = b[self.i:j].strip()
            self.i = j + 1
            if ln:
                return ln
        return None

    def peek_line(self):
        save = self.i
        ln = self.line()
        self.i = save
        return ln

    def binary(self, dtype, count):
        dt = np.dtype(dtype)
        nbytes = dt.itemsize * count
        if self.i + nbytes > len(self.b):
            raise VTKFormatError("truncated binary block")
        a = np.frombuffer(self.b, dtype=dt, count=count, offset=self.i)
        self.i += nbytes
        return a

    def ascii(self, dtype, count):
        if count == 0:
            return np.empty(0, dtype=np.dtype(dtype).newbyteorder("="))
        native = _native()
        if native is not None:
            integer = np.dtype(dtype).kind != "f"
            try:
                vals, used = native.parse_ascii(self.b, self.i, count, integer=integer)
            except native.TractGeomError as e:
                raise VTKFormatError(str(e)) from e
            self.i += used
            return vals if integer else vals.astype(np.dtype(dtype).newbyteorder("="), copy=False)
        # split() is C speed; the remainder stays unparsed
        parts = self.b[self.i:].split(None, count)
        if len(parts) < count:
            raise VTKFormatError("truncated ASCII block")
        rest = parts[count] if len(parts) > count else b""
        self.i = len(self.b) - len(rest)
        kind = np.dtype(dtype).kind
        if kind == "f":
            return np.array(parts[:count], dtype=np.float64).astype(np.dtype(dtype).newbyteorder("="), copy=False)
        return np.array([int(p) for p in parts[:count]], dtype=np.int64)


def _read_bytes(path):
    path = os.fspath(path)
    with open(path, "rb") as f:
        head = f.read(2)
        f.seek(0)
        if head == b"\x1f\x8b":
            return gzip.decompress(f.read())
        return f.read()


def _parse_header(cur):
    """Magic, title, ASCII|BINARY, DATASET POLYDATA -> True when the file is binary."""
    buf = cur.b
    magic = cur.line()
    if magic is None or not magic.lower().startswith(b"# vtk datafile"):
        raise VTKFormatError("not a legacy VTK file")
    # title line may be blank in principle; take the raw next line
    j = buf.find(b"\n", cur.i)
    cur.i = (j + 1) if j >= 0 else len(buf)
    fmt = cur.line()
    if fmt is None or fmt.upper() not in (b"ASCII", b"BINARY"):
        raise VTKFormatError(f"expected ASCII or BINARY, got {fmt!r}")
    ds = cur.line()
    if ds is None or ds.upper().split() != [b"DATASET", b"POLYDATA"]:
        raise VTKFormatError(f"only DATASET POLYDATA is supported, got {ds!r}")
    return fmt.upper() == b"BINARY"


def _skip_field(cur, tok, binary):
    """FIELD <name> <numArrays>, then per array `<name> <numComponents> <numTuples> <type>` + data (vtkPolyDataWriter
    puts such a block BEFORE the geometry when the data set carries field data)."""
    try:
        n_arrays = int(tok[2])
    except (IndexError, ValueError) as e:
        raise VTKFormatError("malformed FIELD header") from e
    for _ in range(n_arrays):
        ln = cur.line()
        if ln is None:
            raise VTKFormatError("truncated FIELD block")
        t = ln.split()
        try:
            comps, tuples, typ = int(t[1]), int(t[2]), t[3].lower()
        except (IndexError, ValueError) as e:
            raise VTKFormatError(f"malformed FIELD array header {ln[:40]!r}") from e
        count = comps * tuples
        if typ == b"string":
            for _ in range(count):                       # one (binary: length-prefixed-free, newline-terminated) line per string
                j = cur.b.find(b"\n", cur.i)
                cur.i = (j + 1) if j >= 0 else len(cur.b)
            continue
        if typ == b"bit":
            if binary:
                cur.i += (count + 7) // 8
            else:
                cur.ascii(">i4", count)
            continue
        dt = _VTK_TYPES.get(typ)
        if dt is None:
            raise VTKFormatError(f"unknown FIELD array type {t[3]!r}")
        if binary:
            cur.binary(dt, count)
        else:
            cur.ascii(dt if np.dtype(dt).kind == "f" else ">i8", count)


def _parse_sections(cur, binary, points=None, keep_order=False, lazy_identity=False):
    """Sections after the header -> (points (P,3), offsets int64[S+1], connectivity int64[C]).

    ``points`` given: the POINTS block was already consumed by the caller (streaming reader).  ``keep_order``: binary
    points stay in the file's big-endian byte order (a zero-copy view; the device swaps them), else native-endian."""
    buf = cur.b
    offsets = None
    conn = None
    while True:
        ln = cur.line()
        if ln is None:
            break
        tok = ln.split()
        key = tok[0].upper()
        if key == b"POINTS":
            n = int(tok[1])
            dt = _VTK_TYPES.get(tok[2].lower())
            if dt is None:
                raise VTKFormatError(f"unknown POINTS type {tok[2]!r}")
            flat = cur.binary(dt, 3 * n) if binary else cur.ascii(dt, 3 * n)
            if flat.dtype.kind != "f":
                flat = flat.astype(np.float64)
            if not (keep_order and binary):
                flat = flat.astype(flat.dtype.newbyteorder("="), copy=False)
            points = flat.reshape(n, 3)
        elif key == b"METADATA":
            # INFORMATION block, terminated by a blank line
            while True:
                j = buf.find(b"\n", cur.i)
                if j < 0:
                    cur.i = len(buf)
                    break
                blank = not buf[cur.i:j].strip()
                cur.i = j + 1
                if blank:
                    break
        elif key in _CELL_SECTIONS:
            a, b = int(tok[1]), int(tok[2])
            nxt = cur.peek_line()
            if nxt is not None and nxt.split()[0].upper() == b"OFFSETS":      # VTK >= 5.1
                t = cur.line().split()
                odt = _VTK_TYPES.get(t[1].lower(), ">i8")
                off = cur.binary(odt, a) if binary else cur.ascii(odt, a)
                t = cur.line().split()
                if t[0].upper() != b"CONNECTIVITY":
                    raise VTKFormatError("expected CONNECTIVITY")
                cdt = _VTK_TYPES.get(t[1].lower(), ">i8")
                cn = cur.binary(cdt, b) if binary else cur.ascii(cdt, b)
                if key == b"LINES":
                    offsets = off.astype(np.int64)
                    conn = cn.astype(np.int64)
            else:                                                            # classic: a cells, b ints
                flat = cur.binary(">i4", b) if binary else cur.ascii(">i4", b)
                if key == b"LINES":
                    native = _native() if binary else None
                    if native is not None:                                   # one native pass over the file's own bytes
                        try:
                            offsets, conn = native.vtk_cells_be32_to_csr(flat)
                        except native.TractGeomError as e:
                            raise VTKFormatError("corrupt LINES array") from e
                    else:
                        offsets, conn = legacy_lines_to_csr(flat.astype(np.int64), a)
        elif key == b"FIELD":
            if points is not None and offsets is not None:
                break          # geometry complete; attributes are not used by this path
            _skip_field(cur, tok, binary)                                    # field data in front of the geometry
        elif key in (b"POINT_DATA", b"CELL_DATA"):
            break              # attributes follow the geometry and are not used by this path
        else:
            raise VTKFormatError(f"unsupported section {ln[:40]!r}")
    if points is None:
        raise VTKFormatError("no POINTS section")
    if offsets is None:
        offsets = np.zeros(1, dtype=np.int64)
        conn = np.empty(0, dtype=np.int64)
    if conn is None and not lazy_identity:                 # the native cell walk reports an identity connectivity as None
        conn = np.arange(int(offsets[-1]), dtype=np.int64)
    return points, offsets, conn


def parse_legacy_polydata(buf: bytes, keep_order=False, lazy_identity=False):
    """-> (points (P,3) float32/float64 — native-endian, or the file's big-endian order with ``keep_order`` —,
    offsets int64[S+1], connectivity int64[C])."""
    cur = _Cursor(buf)
    binary = _parse_header(cur)
    return _parse_sections(cur, binary, keep_order=keep_order, lazy_identity=lazy_identity)


def legacy_lines_to_csr(lines, n_cells=None):
    """Legacy ``[n, i0..i(n-1), n, ...]`` -> (offsets int64[S+1], connectivity int64[C]).

    This is the walk at tract_geom_proc.py:17-25, done with a doubling scan instead of a Python
    ``while``: O(log S) vectorised passes when n_cells is unknown, one pass when it is known and
    every cell has the same size, otherwise a short loop over cells in numpy chunks.
    """
    lines = np.asarray(lines, dtype=np.int64)
    if lines.size == 0:
        return np.zeros(1, dtype=np.int64), np.empty(0, dtype=np.int64)
    # head positions: h0 = 0, h_{i+1} = h_i + 1 + lines[h_i].  Sequential by nature; walk it with a
    # tight loop over the (small) head array only — S iterations of integer work, no point data.
    heads = []
    i, L = 0, lines.size
    if n_cells is not None and n_cells > 0:
        # fast path: uniform cell size
        n0 = int(lines[0])
        if (1 + n0) * n_cells == L:
            h = np.arange(n_cells, dtype=np.int64) * (1 + n0)
            if np.all(lines[h] == n0):
                offsets = np.arange(n_cells + 1, dtype=np.int64) * n0
                mask = np.ones(L, dtype=bool); mask[h] = False
                return offsets, lines[mask]
    native = _native()
    if native is not None:
        try:
            return native.vtk_lines_to_csr(lines)
        except native.TractGeomError as e:
            raise VTKFormatError("corrupt LINES array") from e
    lst = lines.tolist() if L < (1 << 26) else lines
    while i < L:
        n = int(lst[i])
        if n < 0 or i + 1 + n > L:                    # checked INSIDE the walk: a negative count would never advance
            raise VTKFormatError("corrupt LINES array")
        heads.append(i)
        i += 1 + n
    h = np.asarray(heads, dtype=np.int64)
    counts = lines[h]
    if counts.min(initial=0) < 0 or (h[-1] + 1 + counts[-1]) > L:
        raise VTKFormatError("corrupt LINES array")
    offsets = np.zeros(len(h) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    mask = np.ones(L, dtype=bool); mask[h] = False
    return offsets, lines[mask]


def _apply_connectivity(pts, off, conn):
    if conn is None:                                        # the native cell walk found the identity
        if int(off[-1]) != len(pts):
            if int(off[-1]) > len(pts):
                raise VTKFormatError("LINES connectivity refers to a point that does not exist")
            return pts[:int(off[-1])]
        return pts
    if conn.size and (conn.min() < 0 or conn.max() >= len(pts)):
        raise VTKFormatError("LINES connectivity refers to a point that does not exist")
    identity = conn.size == len(pts) and (conn.size == 0 or (conn[0] == 0 and conn[-1] == conn.size - 1 and np.all(np.diff(conn) == 1)))
    return pts if identity else pts[conn]                   # tract_geom_proc.py:19-20 in one gather


def read_polylines_csr(path, dtype=None):
    """File -> (points_csr (C,3), offsets int64[S+1]) with the connectivity already applied, native byte order.

    ``dtype=None`` keeps the file's point dtype (float32 for the usual ``POINTS n float``);
    pass np.float64 for the canonical parity input (exact upcast).
    """
    pts, off, conn = _read_any(path)
    out = _apply_connectivity(pts, off, conn)
    if dtype is not None:
        out = out.astype(dtype, copy=False)
    return np.ascontiguousarray(out), off


def read_polylines_raw(path, arena=None):
    """File -> (points (C,3), offsets int64[S+1]) for the device path: no per-point work on the host.

    Binary legacy files keep their big-endian ``float``/``double`` bytes (the device swaps and upcasts them,
    tg_dtype TG_F32_BE / TG_F64_BE).  With ``arena`` (a ``_lib.PinnedArena``) the POINTS block of an uncompressed
    file is read from the file STRAIGHT into pinned memory (``readinto``; a .vtk.gz block is copied there after
    the in-memory gunzip), so the host-to-device copy that follows is an asynchronous DMA transfer."""
    p = os.fspath(path)
    low = p.lower()
    if not (low.endswith(".vtk") or low.endswith(".vtk.gz")):
        return read_polylines_csr(p)
    if not os.path.exists(p):
        raise FileNotFoundError(p)
    with open(p, "rb") as f:
        head = f.read(1 << 16)
        hit = None if head[:2] == b"\x1f\x8b" else _leading_points_block(head)
        if hit is None:                                     # gzip, ASCII, or something in front of POINTS: whole-buffer parse
            buf = head + f.read()
            if buf[:2] == b"\x1f\x8b":
                buf = gzip.decompress(buf)
            pts, off, conn = parse_legacy_polydata(buf, keep_order=True, lazy_identity=True)
            out = _apply_connectivity(pts, off, conn)
            if arena is not None and out.size:
                pinned = arena.take(out.nbytes).view(out.dtype).reshape(out.shape)
                np.copyto(pinned, out)
                out = pinned
            return np.ascontiguousarray(out), off
        n, dt, start = hit
        nbytes = 3 * n * np.dtype(dt).itemsize
        block = arena.take(nbytes) if arena is not None else np.empty(nbytes, dtype=np.uint8)
        have = min(len(head) - start, nbytes)
        block[:have] = np.frombuffer(head, dtype=np.uint8, count=have, offset=start)
        if have < nbytes:
            got = f.readinto(memoryview(block)[have:])
            if got != nbytes - have:
                raise VTKFormatError("truncated binary block")
            rest = f.read()
        else:
            rest = head[start + nbytes:] + f.read()
    pts = block.view(dt).reshape(n, 3)
    _, off, conn = _parse_sections(_Cursor(rest), True, points=pts, keep_order=True, lazy_identity=True)
    out = _apply_connectivity(pts, off, conn)
    return (out if out.flags.c_contiguous else np.ascontiguousarray(out)), off


def _leading_points_block(head: bytes):
    """(n, big-endian dtype, byte offset of the data) when ``head`` starts a BINARY POLYDATA file whose first section is
    POINTS with a float type; None otherwise (the caller then parses the whole buffer)."""
    try:
        cur = _Cursor(head)
        if not _parse_header(cur):
            return None
        ln = cur.line()
        if ln is None:
            return None
        tok = ln.split()
        if tok[0].upper() != b"POINTS":
            return None
        dt = _VTK_TYPES.get(tok[2].lower())
        if dt is None or np.dtype(dt).kind != "f" or cur.i > len(head):
            return None
        return int(tok[1]), dt, cur.i
    except (VTKFormatError, IndexError, ValueError):
        return None


def _read_any(path):
    p = os.fspath(path)
    low = p.lower()
    if low.endswith(".vtk") or low.endswith(".vtk.gz"):
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        return parse_legacy_polydata(_read_bytes(p), lazy_identity=True)
    try:
        import pyvista as pv   # optional: other formats
    except ImportError as e:
        raise VTKFormatError(f"{p}: only legacy .vtk/.vtk.gz can be read without pyvista") from e
    mesh = pv.read(p)
    off, conn = legacy_lines_to_csr(np.asarray(mesh.lines))
    return np.asarray(mesh.points), off, conn


# ---------------------------------------------------------------------------------------------
# Writer (tests, synthetic fixtures, bench end-to-end)
# ---------------------------------------------------------------------------------------------
def write_polylines(path, points, offsets, binary=True, point_dtype="double", layout="classic", connectivity=None,
                    title="tractgeom synthetic"):
    """Write a legacy POLYDATA file.  ``layout`` is 'classic' (LINES n size) or 'offsets' (5.1)."""
    points = np.asarray(points).reshape(-1, 3)
    offsets = np.asarray(offsets, dtype=np.int64)
    S = len(offsets) - 1
    if connectivity is None:
        connectivity = np.arange(int(offsets[-1]), dtype=np.int64)
    connectivity = np.asarray(connectivity, dtype=np.int64)
    fdt = {"float": np.float32, "double": np.float64}[point_dtype]
    ver = b"5.1" if layout == "offsets" else b"3.0"
    chunks = [b"# vtk DataFile Version " + ver + b"\n", title.encode()[:255] + b"\n",
              b"BINARY\n" if binary else b"ASCII\n", b"DATASET POLYDATA\n",
              f"POINTS {len(points)} {point_dtype}\n".encode()]
    if binary:
        chunks.append(points.astype(np.dtype(fdt).newbyteorder(">")).tobytes())
        chunks.append(b"\n")
    else:
        fmt = "%.9g" if point_dtype == "float" else "%.17g"
        chunks.append("\n".join(" ".join(fmt % v for v in row) for row in points.astype(fdt)).encode() + b"\n")
    if layout == "classic":
        n = np.diff(offsets)
        flat = np.empty(len(connectivity) + S, dtype=np.int64)
        head = offsets[:-1] + np.arange(S, dtype=np.int64)
        mask = np.ones(len(flat), dtype=bool)
        if S:
            flat[head] = n
            mask[head] = False
        flat[mask] = connectivity
        chunks.append(f"LINES {S} {len(flat)}\n".encode())
        if binary:
            chunks.append(flat.astype(">i4").tobytes()); chunks.append(b"\n")
        else:
            chunks.append(" ".join(map(str, flat.tolist())).encode() + b"\n")
    elif layout == "offsets":
        chunks.append(f"LINES {S + 1} {len(connectivity)}\n".encode())
        chunks.append(b"OFFSETS vtktypeint64\n")
        if binary:
            chunks.append(offsets.astype(">i8").tobytes()); chunks.append(b"\n")
        else:
            chunks.append(" ".join(map(str, offsets.tolist())).encode() + b"\n")
        chunks.append(b"CONNECTIVITY vtktypeint64\n")
        if binary:
            chunks.append(connectivity.astype(">i8").tobytes()); chunks.append(b"\n")
        else:
            chunks.append(" ".join(map(str, connectivity.tolist())).encode() + b"\n")
    else:
        raise ValueError(layout)
    data = b"".join(chunks)
    p = os.fspath(path)
    if p.lower().endswith(".gz"):
        with gzip.open(p, "wb", compresslevel=1) as f:
            f.write(data)
    else:
        with open(p, "wb") as f:
            f.write(data)
    return p
