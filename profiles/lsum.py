"""Compact per-launch view of an ncu --csv launch list: python profiles/lsum.py launches.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    k = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]])
    agg.setdefault(k, {})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
tot = 0
for k, m in agg.items():
    t = m["gpu__time_duration.sum"] / 1e3; tot += t
    print("%-22s %-16s %7.1f us  rd %6.1f MB wr %6.1f MB  fp64 %4.1f%%  issue %4.1f%%" % (k[1][:22], k[2], t, m.get("dram__bytes_read.sum", 0) / 1e6,
          m.get("dram__bytes_write.sum", 0) / 1e6, m.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0), m.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0)))
print("total %.1f us" % tot)
