#!/usr/bin/env python
"""One text summary of an .ncu-rep for profiles/: headline counters, warp-stall reasons, local-memory traffic, hottest
source lines.  usage: ncu_report.py report.ncu-rep "title / command line" > profiles/rNN_ncu_<what>.txt"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
print("# " + title)
print(subprocess.run([sys.executable, __file__.replace("ncu_report.py", "ncu_summary.py"), rep], capture_output=True,
                     text=True).stdout.strip())
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
st = []
for h, v in zip(hdr, vals):
    if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
        try:
            st.append((float(v.replace(",", "")), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
tot = sum(v for v, _ in st) or 1.0
print("\n== warp stall reasons (share of the sampled warp states)")
for v, h in sorted(st, reverse=True)[:10]:
    print("%-28s %5.1f %%" % (h, 100.0 * v / tot))
print("\n== local memory (spills) and shared memory")
for h, v in zip(hdr, vals):
    if h in ("l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
             "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_active.avg"):
        print("%-60s %s" % (h, v))
print("\n== hottest source lines")
print(subprocess.run([sys.executable, __file__.replace("ncu_report.py", "ncu_lines.py"), rep, "0", "18"],
                     capture_output=True, text=True).stdout.strip())
