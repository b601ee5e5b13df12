#!/usr/bin/env python
"""SASS-level view of an .ncu-rep (needs --import-source on): the instructions that collect the most warp-stall
samples with their two main stall reasons, and the sample / stall-reason share of address ranges given as
name=first:last (instruction indices as printed in the first column).
usage: ncu_sass.py report.ncu-rep [top-N] [name=first:last ...]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
regions = [a for a in sys.argv[3:] if "=" in a]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def smp(r):
    return int(r[ix["# Samples"]] or 0)


def main_stalls(rs, k):
    tot = sum(smp(r) for r in rs) or 1
    d = sorted(((sum(int(r[ix[h]] or 0) for r in rs), h[6:]) for h in stalls), reverse=True)[:k]
    return ", ".join("%s %.1f %%" % (n, 100.0 * v / tot) for v, n in d if v)


tot = sum(smp(r) for r in data) or 1
print("total samples %d over %d instructions" % (tot, len(data)))
for i in sorted(sorted(range(len(data)), key=lambda i: -smp(data[i]))[:top]):
    r = data[i]
    print("%6d %6.2f %%  x%-11s %-58s %s" % (i, 100.0 * smp(r) / tot, r[ix["Instructions Executed"]],
                                            r[1].strip()[:58], main_stalls([r], 2)))
for reg in regions:
    name, _, span = reg.partition("=")
    a, b = (int(v) for v in span.split(":"))
    rs = data[a:b + 1]
    s = sum(smp(r) for r in rs)
    x = sum(int(r[ix["Instructions Executed"]] or 0) for r in rs)
    print("== %-28s %6.2f %% of the samples, %.3e warp instructions: %s" % (name, 100.0 * s / tot, x, main_stalls(rs, 6)))
