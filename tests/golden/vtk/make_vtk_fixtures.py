#!/usr/bin/env python3
"""Byte-level legacy-VTK fixtures in the layouts vtkPolyDataWriter emits, assembled here with
`struct` — deliberately NOT with lesion_condition_vae_b200.vtk_io.write_polylines, so that the reader is checked
against an independent statement of the format (vtkDataWriter.cxx / vtkPolyDataWriter.cxx, VTK 9):

  v42_field_float.vtk   "# vtk DataFile Version 4.2", BINARY, a FIELD FieldData block (two arrays, one of them a string
                        array) IN FRONT of the geometry, POINTS n float, classic LINES n size (int32), then CELL_DATA
  v51_offsets.vtk       "Version 5.1", BINARY, POINTS n double, METADATA/INFORMATION block, LINES with
                        OFFSETS vtktypeint64 / CONNECTIVITY vtktypeint64, POINT_DATA with a scalar array
  v51_offsets.vtk.gz    the same bytes, gzip (what <tract>_curves.vtk.gz holds,
                        /root/reference/src/geometry/comprehensive_tract_geometry_analysis.py:86)
  v30_ascii_field.vtk   ASCII, FIELD before POINTS, float points, shuffled connectivity (not the identity)
  expected.npz          points (float64, exact upcast of what the file stores) / offsets / connectivity-applied CSR

usage: python tests/golden/vtk/make_vtk_fixtures.py   (writes next to itself)"""
import gzip
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def be(fmt, values):
    return struct.pack(">" + fmt * len(values), *values)


def polylines(seed, lengths):
    rng = np.random.default_rng(seed)
    pts, off = [], [0]
    for n in lengths:
        p = np.cumsum(rng.normal(size=(n, 3)) * 0.4 + 0.3, axis=0) + rng.uniform(-40, 40, 3)
        pts.append(p)
        off.append(off[-1] + n)
    return np.concatenate(pts), np.asarray(off, dtype=np.int64)


def main():
    expected = {}
    # ---- v4.2, FIELD in front, float points, classic cells ---------------------------------------------------
    pts, off = polylines(1, [5, 3, 12, 2, 7])
    p32 = pts.astype(np.float32)
    cells = []
    for s in range(len(off) - 1):
        cells += [int(off[s + 1] - off[s])] + list(range(int(off[s]), int(off[s + 1])))
    b = b"# vtk DataFile Version 4.2\nvtk output\nBINARY\nDATASET POLYDATA\n"
    b += b"FIELD FieldData 2\n"
    b += b"subject_age 1 3 double\n" + be("d", [61.5, 9.0, 2.0]) + b"\n"
    b += b"note 1 1 string\nbundle%20of%20tests\n"
    b += f"POINTS {len(p32)} float\n".encode() + be("f", p32.reshape(-1).tolist()) + b"\n"
    b += f"LINES {len(off) - 1} {len(cells)}\n".encode() + be("i", cells) + b"\n"
    b += f"CELL_DATA {len(off) - 1}\nSCALARS id int 1\nLOOKUP_TABLE default\n".encode() + be("i", list(range(len(off) - 1))) + b"\n"
    open(os.path.join(HERE, "v42_field_float.vtk"), "wb").write(b)
    expected["v42_field_float/points"], expected["v42_field_float/offsets"] = p32.astype(np.float64), off

    # ---- v5.1, double points, METADATA, OFFSETS / CONNECTIVITY int64 -------------------------------------------
    pts, off = polylines(2, [4, 40, 3, 9])
    b = b"# vtk DataFile Version 5.1\nvtk output\nBINARY\nDATASET POLYDATA\n"
    b += f"POINTS {len(pts)} double\n".encode() + be("d", pts.reshape(-1).tolist()) + b"\n"
    b += b"METADATA\nINFORMATION 0\n\n"
    b += f"LINES {len(off)} {int(off[-1])}\n".encode()
    b += b"OFFSETS vtktypeint64\n" + be("q", off.tolist()) + b"\n"
    b += b"CONNECTIVITY vtktypeint64\n" + be("q", list(range(int(off[-1])))) + b"\n"
    b += f"POINT_DATA {len(pts)}\nSCALARS fa float\nLOOKUP_TABLE default\n".encode() + be("f", np.linspace(0, 1, len(pts)).tolist()) + b"\n"
    open(os.path.join(HERE, "v51_offsets.vtk"), "wb").write(b)
    with open(os.path.join(HERE, "v51_offsets.vtk.gz"), "wb") as f:
        f.write(gzip.compress(b, compresslevel=6, mtime=0))
    expected["v51_offsets/points"], expected["v51_offsets/offsets"] = pts, off

    # ---- ASCII, FIELD in front, connectivity that is not the identity ---------------------------------------
    pts, off = polylines(3, [6, 3, 5])
    p32 = pts.astype(np.float32)
    perm = np.random.default_rng(4).permutation(len(p32))           # file order of the points
    inv = np.argsort(perm)                                          # point k of the CSR is file point inv[k]
    filed = p32[perm]
    lines = []
    for s in range(len(off) - 1):
        ids = inv[int(off[s]):int(off[s + 1])]
        lines.append(" ".join([str(len(ids))] + [str(int(i)) for i in ids]))
    t = "# vtk DataFile Version 3.0\nascii with field\nASCII\nDATASET POLYDATA\n"
    t += "FIELD FieldData 1\nweights 2 2 float\n0.5 1.5\n2.5 3.5\n"
    t += f"POINTS {len(filed)} float\n" + "\n".join(" ".join(repr(float(v)) for v in row) for row in filed) + "\n"
    t += f"LINES {len(off) - 1} {len(off) - 1 + int(off[-1])}\n" + "\n".join(lines) + "\n"
    open(os.path.join(HERE, "v30_ascii_field.vtk"), "w").write(t)
    expected["v30_ascii_field/points"], expected["v30_ascii_field/offsets"] = p32.astype(np.float64), off
    np.savez(os.path.join(HERE, "expected.npz"), **expected)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
