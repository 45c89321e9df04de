#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv` export by address region.
usage: python tools/ncu_regions.py src.csv name=lo:hi [name=lo:hi ...]   (hex offsets relative to the kernel start)"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][ix["Address"]], 16)
regs = []
for a in sys.argv[2:]:
    nm, r = a.split("=")
    lo, hi = r.split(":")
    regs.append((nm, int(lo, 16), int(hi, 16)))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print(f"total: {tot_inst} warp-instructions, {tot_samp} samples")
def summarize(nm, sel):
    inst = sum(int(r[ix["Instructions Executed"]]) for r in sel)
    samp = sum(int(r[ix["# Samples"]]) for r in sel)
    st = {s: sum(int(r[ix[s]]) for r in sel) for s in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:6]
    print(f"{nm:14s} static {len(sel):5d}  executed {inst:12d} ({100*inst/tot_inst:5.1f}%)  samples {samp:8d} ({100*samp/tot_samp:5.1f}%)  samples/inst {samp/max(inst,1)*1e3:7.3f}e-3  " +
          " ".join(f"{k[6:]}={100*v/max(samp,1):.0f}%" for k, v in top))
covered = set()
for nm, lo, hi in regs:
    sel = [r for r in data if lo <= int(r[ix["Address"]], 16) - base <= hi]
    covered.update(id(r) for r in sel)
    summarize(nm, sel)
summarize("(rest)", [r for r in data if id(r) not in covered])
