"""per-kernel summary of an ncu --csv launch list: launches, mean duration, mean DRAM bytes read/written"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ix = {h: k for k, h in enumerate(hdr)}
acc = collections.OrderedDict()
for r in rows[1:]:
    name = r[ix["Kernel Name"]].split("(")[0][:60]; m = r[ix["Metric Name"]]; u = r[ix["Metric Unit"]]; v = float(r[ix["Metric Value"]].replace(",", ""))
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}.get(u, 1.0)
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    acc.setdefault(name, collections.defaultdict(list))[m].append(v)
print("%-60s %6s %10s %10s %10s" % ("kernel", "n", "ms(mean)", "rd MB", "wr MB"))
for name, d in sorted(acc.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    t = d["gpu__time_duration.sum"]; rd = d.get("dram__bytes_read.sum", [0]); wr = d.get("dram__bytes_write.sum", [0])
    print("%-60s %6d %10.4f %10.1f %10.1f" % (name, len(t), sum(t) / len(t), sum(rd) / len(rd) / 1e6, sum(wr) / len(wr) / 1e6))
