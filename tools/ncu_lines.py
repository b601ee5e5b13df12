#!/usr/bin/env python
"""Per-source-line stall samples and executed instructions from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [kernel-index] [top-N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# split per kernel: each starts with a "File Path" row
starts = [i for i, r in enumerate(rows) if r and r[0] == "File Path"]
starts.append(len(rows))
# one block per (source file, kernel); kidx selects the kernel (by order of first appearance)
kernels = []
for a in starts[:-1]:
    if rows[a + 1][1] not in kernels:
        kernels.append(rows[a + 1][1])
kname = kernels[kidx]
print(kname)
lines = []
for a, b in zip(starts[:-1], starts[1:]):
    blk = rows[a:b]
    if blk[1][1] != kname:
        continue
    fname = blk[0][1].split("/")[-1]
    hdr = blk[2]
    i_s = hdr.index("# Samples"); i_x = hdr.index("Instructions Executed")
    for r in blk[3:]:
        if r[0] != "":
            try:
                lines.append(("%s:%s" % (fname, r[0]), r[1], int(r[i_s] or 0), int(r[i_x] or 0)))
            except ValueError:
                pass
tot_s = sum(l[2] for l in lines) or 1
tot_x = sum(l[3] for l in lines) or 1
print("total samples %d, total warp instructions %d" % (tot_s, tot_x))
for ln, src, s, x in sorted(lines, key=lambda l: -l[2])[:top]:
    print("%-22s %6.2f%% smp %6.2f%% inst  %s" % (ln, 100.0 * s / tot_s, 100.0 * x / tot_x, src.strip()[:110]))
