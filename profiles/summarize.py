"""Turn an `ncu --csv --log-file` launch list (gpu__time_duration.sum, dram bytes) into a markdown table.

    python profiles/summarize.py gpurun_out/launches.csv > profiles/launches_rNN_<op>.md
"""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    agg, order = collections.defaultdict(dict), []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        k = (int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Grid Size"]], r[ix["Block Size"]])
        if k not in agg:
            order.append(k)
        v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "byte": v / 1e6, "Kbyte": v / 1e3, "Mbyte": v, "Gbyte": v * 1e3}.get(u, v)
        agg[k][r[ix["Metric Name"]]] = v
    tot = sum(agg[k]["gpu__time_duration.sum"] for k in order)
    by = collections.Counter()
    print("| # | kernel | grid | block | time (us) | share | dram read (MB) | dram write (MB) |")
    print("|---|---|---|---|---|---|---|---|")
    for n, k in enumerate(order):
        m = agg[k]
        t = m["gpu__time_duration.sum"]
        by[k[1]] += t
        print("| %d | `%s` | %s | %s | %.1f | %.1f%% | %.1f | %.1f |" % (n, k[1], k[2], k[3], t, 100 * t / tot,
                                                                   m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)))
    print("\ntotal %.1f us over %d launches (cold-cache, serialised by ncu: compare shares, not absolutes)\n" % (tot, len(order)))
    print("| kernel | time (us) | share |\n|---|---|---|")
    for k, t in by.most_common():
        print("| `%s` | %.1f | %.1f%% |" % (k, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1])
