#!/usr/bin/env python3
"""Basic-block view of an `ncu --page source --csv` export: contiguous runs of instructions with the
same execution count, with their share of executed instructions and of stall samples.
usage: python tools/ncu_blocks.py src.csv [min_sample_pct]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
blocks = []
cur = None
for r in data:
    c = int(r[ix["Instructions Executed"]])
    if cur is None or c != cur[0]:
        cur = [c, []]; blocks.append(cur)
    cur[1].append(r)
print(f"total {tot_i} warp-instr, {tot_s} samples")
for c, rs in blocks:
    samp = sum(int(r[ix["# Samples"]]) for r in rs)
    if 100 * samp / tot_s < minpct: continue
    a0 = int(rs[0][ix["Address"]], 16) - base; a1 = int(rs[-1][ix["Address"]], 16) - base
    ops = collections.Counter(r[ix["Source"]].split()[0 if not r[ix["Source"]].strip().startswith("@") else 1].split(".")[0] for r in rs)
    f64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    st = {s: sum(int(r[ix[s]]) for r in rs) for s in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    print(f"0x{a0:05x}-0x{a1:05x} n={len(rs):4d} exec/instr={c:9d} instr%={100*c*len(rs)/tot_i:5.1f} samp%={100*samp/tot_s:5.1f} fp64={f64:3d} " +
          " ".join(f"{k[6:]}={100*v/max(samp,1):.0f}%" for k, v in top))
