#!/usr/bin/env python
"""Attribute the warp-stall samples / executed instructions of an ncu --set full capture to CUDA source
lines: joins `ncu --page source --print-source sass` with `nvdisasm -g` line info by instruction order.
usage: tools/ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [launch-index]"""
import csv, io, re, subprocess, sys, os, tempfile, collections
rep, so, kern = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name": cur = {"name": row[1], "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None or not row: continue
    if cur["hdr"] is None: cur["hdr"] = row; continue
    cur["rows"].append(row)
b = blocks[which]
h = b["hdr"]; iS = h.index("# Samples"); iI = h.index("Instructions Executed"); iT = h.index("Thread Instructions Executed")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = []
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin"):
        dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
lines, insec, curline = [], False, None
for l in dis:
    if l.startswith("\t.section"):
        insec = (".text." in l and kern in l); continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        inl = re.findall(r'inlined at "([^"]+)", line (\d+)', l)
        curline = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(c)) for a, c in inl)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): lines.append(curline)
rows = b["rows"]
print("kernel:", b["name"], " sass rows:", len(rows), " disasm insts:", len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for k, r in enumerate(rows):
    key = lines[k] if k < len(lines) and lines[k] else ("?", 0, ())
    # attribute to the outermost frame inside the kernel file + innermost line
    v = (int(r[iS] or 0), int(r[iI] or 0), int(r[iT] or 0))
    for q in range(3): agg[key][q] += v[q]; tot[q] += v[q]
print("total samples %d, warp insts %d, thread insts %d" % tuple(tot))
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(os.environ.get("TOP", "60"))]
for (f, ln, inl), v in top:
    print("%5.1f%% smp %5.1f%% inst  thr/inst %4.1f  %s:%d %s" % (100.0 * v[0] / tot[0], 100.0 * v[1] / tot[1], v[2] / max(v[1], 1), f, ln, " <- ".join("%s:%d" % x for x in inl)))
