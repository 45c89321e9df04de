#!/usr/bin/env python3
"""ncu `--csv --log-file` launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per
launch, long format) -> the text table kept under profiles/.
usage: python tools/launch_list.py launches.csv > profiles/rN_launches.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
launch = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["ID"]] == "ID":
        continue
    d = launch.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]].split("(")[0].split("::")[-1]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    name = r[ix["Metric Name"]]
    if name.startswith("gpu__time"):
        d["ms"] = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    else:
        mb = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1e-6)
        d["rd" if "read" in name else "wr"] = mb
print(f"{'id':>3} {'kernel':<24} {'ms':>9} {'dram rd MB':>12} {'dram wr MB':>12}")
for i, d in launch.items():
    print(f"{i:3d} {d['name']:<24} {d.get('ms', 0):9.4f} {d.get('rd', 0):12.1f} {d.get('wr', 0):12.1f}")
