#!/usr/bin/env python
"""Dynamic instruction mix + stall-reason shares of one kernel of an ncu --set full capture (source page, SASS view).
usage: tools/ncu_mix.py <report.ncu-rep> [launch-index]"""
import csv, io, re, subprocess, sys, collections
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name": cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None or not row: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
b = blocks[which]; hdr = b["hdr"]
iS = hdr.index("Source"); iI = hdr.index("Instructions Executed"); iSm = hdr.index("# Samples")
stall = [(n, h) for n, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
ops, smp, st, tot = collections.Counter(), collections.Counter(), collections.Counter(), 0
for r in b["rows"]:
    if len(r) < len(hdr) - 2: continue
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', r[iS]); op = m.group(2) if m else "?"
    n = int(r[iI] or 0); ops[op] += n; tot += n; smp[op] += int(r[iSm] or 0)
    for c, h in stall: st[h] += int(r[c] or 0)
print("kernel:", b["name"], " static", len(b["rows"]), " warp instructions", tot)
ss = sum(smp.values())
for k, v in ops.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30): print("%-9s %6.2f%% inst  %6.2f%% samples" % (k, 100 * v / tot, 100 * smp[k] / ss))
print({k: round(100 * v / sum(st.values()), 1) for k, v in st.most_common(10)})
