"""Build libtractgeom.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the .so is
git-ignored but travels with the repo snapshot to the GPU box.

Identity: the library carries a BUILD ID = sha256 over every file of csrc/, include/tractgeom.h and
the nvcc command line (flags + defines), compiled in as TG_BUILD_ID and returned by tg_build_id().
`is_stale()` compares the id found in the library file with the id of the sources on disk — no
mtimes — so the binary that is tested and benchmarked is provably the one HEAD's sources produce
(tests/test_host_cpu.py::test_library_is_built_from_these_sources, tests/conftest.py, bench.py).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libtractgeom.so")
SOURCES = ["tg_kernels.cu"]
PUBLIC_HEADER = os.path.join(ROOT, "include", "tractgeom.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "-ldl",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libtractgeom.so")


def dependency_files():
    """Every file whose content ends up in the binary: all of csrc/ (sources and headers) + the public header."""
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".hpp", ".inc"))]
    return files + [PUBLIC_HEADER]


def source_id(defines=()):
    """Build id of the sources on disk (16 hex digits of a sha256 over file names, contents, flags, defines)."""
    h = hashlib.sha256()
    for path in dependency_files():
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as f:
            h.update(f.read())
        h.update(b"\0")
    h.update(" ".join(NVCC_FLAGS).encode() + b"\0")
    h.update(" ".join(sorted(defines)).encode())
    return h.hexdigest()[:16]


ID_MARKER = b"@(#)TG_BUILD_ID="


def library_id(path=None):
    """Build id compiled into an existing library, or None (missing file, or a library without one).
    Read from the file's bytes, not through dlopen: a process that later loads a rebuilt library must
    not already hold a handle on the old one."""
    path = path or LIB
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        blob = f.read()
    i = blob.find(ID_MARKER)
    if i < 0:
        return None
    return blob[i + len(ID_MARKER):i + len(ID_MARKER) + 16].decode(errors="replace")


def is_stale(defines=()):
    return library_id() != source_id(defines)


def build(force=False, verbose=False, out=None, defines=()):
    """Compile the library when its build id differs from the sources' (or `force`).
    `out`/`defines` build a tuning variant (e.g. defines=["TG_WARPS=10"]) next to the default one;
    select it at run time with the TG_LIB environment variable."""
    if out is None and not force and not is_stale(defines):
        return LIB
    out = out or LIB
    bid = source_id(defines)
    cmd = ([find_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), f'-DTG_BUILD_ID="{bid}"']
           + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out + ".tmp"] + SOURCES)
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    os.replace(out + ".tmp", out)          # a dlopen'ed old copy keeps its inode; new loads see the new file
    if verbose:
        print(res.stdout + res.stderr)
    return out


def steady_block_stats(path=None):
    """Static instruction mix of k_metrics_grouped's steady block (the innermost loop that processes kSub = 3 interior
    points), read from the SASS of the built library with cuobjdump: {"fp64_pipe_per_point", "instructions_per_point",
    "mufu_per_point"}; None when cuobjdump is missing.  bench.py reports these instead of hand-written constants."""
    import re
    path = path or LIB
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(path):
        return None
    try:
        sass = subprocess.run([exe, "-sass", path], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    ins, cur = [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
        if m and cur and "k_metrics_grouped" in cur:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr = {a: i for i, (a, _) in enumerate(ins)}
    best = None
    for i, (a, t) in enumerate(ins):
        m = re.search(r"\bBRA(?:\.\w+)*\s+0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr:
            body = [x for _, x in ins[addr[tgt]:i + 1]]
            ops = [re.sub(r"^@!?U?P\w+\s+", "", x).split()[0].split(".")[0] for x in body]
            f64 = sum(o in ("DFMA", "DMUL", "DADD", "DSETP") for o in ops)
            if f64 >= 150 and (best is None or len(body) < best[0]):      # the smallest loop that holds a whole 3-point block
                best = (len(body), f64, sum(o == "MUFU" for o in ops))
    if best is None:
        return None
    return {"fp64_pipe_per_point": best[1] / 3.0, "instructions_per_point": best[0] / 3.0, "mufu_per_point": best[2] / 3.0,
            "source": "cuobjdump -sass of the loaded library: innermost loop of k_metrics_grouped over 3 interior points"}


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    if "--id" in sys.argv:
        print("sources:", source_id(defs), "library:", library_id())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=(os.path.join(HERE, outs[0]) if outs else None), defines=defs))
