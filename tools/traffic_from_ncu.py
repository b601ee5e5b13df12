#!/usr/bin/env python
"""DRAM bytes of the fused baseline-selection launch group from an ncu launch list (csv with gpu__time_duration.sum,
dram__bytes_read.sum, dram__bytes_write.sum per launch): -> profiles/r02_traffic_<config>.json, which bench.py quotes
as roofline.traffic when it is run with the same command.
usage: traffic_from_ncu.py launches.csv config genes degnorm_iter kernel_regex out.json"""
import csv
import json
import re
import sys

path, config, genes, n_iter, pattern, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
rows = [r for r in csv.reader(open(path, errors="replace")) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
cols = {c: i for i, c in enumerate(rows[hdr])}
per = {}
for r in rows[hdr + 1:]:
    if len(r) <= cols["Metric Value"]:
        continue
    key = (r[cols["ID"]], r[cols["Kernel Name"]])
    val = float(r[cols["Metric Value"]].replace(",", ""))
    unit = r[cols["Metric Unit"]]
    scale = dict(byte=1.0, Kbyte=1e3, Mbyte=1e6, Gbyte=1e9, Tbyte=1e12, ns=1e-9, us=1e-6, ms=1e-3, s=1.0,
                 nsecond=1e-9, usecond=1e-6, msecond=1e-3, second=1.0).get(unit, 1.0)
    per.setdefault(key, {})[r[cols["Metric Name"]]] = val * scale
sel = [(k, v) for k, v in per.items() if re.search(pattern, k[1])]
tot_t = sum(v.get("gpu__time_duration.sum", 0.0) for v in per.values())
sel_t = sum(v.get("gpu__time_duration.sum", 0.0) for _, v in sel)
dram = sum(v.get("dram__bytes_read.sum", 0.0) + v.get("dram__bytes_write.sum", 0.0) for _, v in sel)
res = dict(config=config, genes=genes, launches_listed=len(per), fused_launches=len(sel), outer_iterations=n_iter,
           dram_bytes_per_launch_group=dram / n_iter, fused_share_of_device_time=sel_t / tot_t if tot_t else None,
           fused_seconds_serialised=sel_t, source=path)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
