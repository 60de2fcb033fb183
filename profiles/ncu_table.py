"""One row per launch of a .ncu-rep: duration, DRAM bytes and rate, pipe utilisation, occupancy, top stall reasons."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
units = rows[1]
SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
def scaled(r, k):
    """value in us (durations) or MB (bytes), whatever unit ncu chose"""
    return g(r, k) * SCALE.get(units[col[k]], 1.0) if k in col else 0.0
def g(r, k, d=0.0):
    try:
        return float(r[col[k]].replace(',', ''))
    except Exception:
        return d
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
print("%-3s %-34s %8s %8s %8s %7s %6s %6s %6s %5s %5s  %s" % ("#", "kernel", "us", "rd MB", "wr MB", "GB/s", "dram%", "fp64%", "issue%", "occ%", "regs", "top stalls"))
for i, r in enumerate(rows[2:]):
    name = r[col['Kernel Name']][:34]
    us = scaled(r, 'gpu__time_duration.sum')
    rd, wr = scaled(r, 'dram__bytes_read.sum'), scaled(r, 'dram__bytes_write.sum')
    st = sorted(((g(r, s), s.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for s in stalls), reverse=True)[:3]
    print("%-3d %-34s %8.1f %8.1f %8.1f %7.0f %6.1f %6.1f %6.1f %5.1f %5.0f  %s" % (
        i, name, us, rd, wr, (rd + wr) / us * 1e3 if us else 0, g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        g(r, 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'), g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'),
        g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'), g(r, 'launch__registers_per_thread'),
        " ".join("%s=%.2f" % (n, v) for v, n in st)))
