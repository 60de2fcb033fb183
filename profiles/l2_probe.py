"""Is the two-pass NTT limited by HBM traffic?  Same launch shape (3 limbs x 37 polynomials = 111 limb-transforms, 55 MB,
sized so that both passes fill the machine evenly), once re-running on ONE buffer (L2-resident after the first run) and
once rotating through 8 buffers (every pass streams from / to HBM)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import homulator_b200 as hml
ctx = hml.Context(os.path.join(ROOT, "config", "config_4.cfg"), 45, 15)
idx = [0, 17, 45]
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 37
bufs = [ctx.uniform(idx, 5 + i, lead=(nb,)) for i in range(8)]
def timeit(fn, reps=24):
    for i in range(4): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
n = nb * len(idx)
for inv in (False, True):
    t_l2 = timeit(lambda i: ctx.ntt_batch(bufs[0], idx, out=bufs[0], inverse=inv))
    t_hbm = timeit(lambda i: ctx.ntt_batch(bufs[i % 8], idx, out=bufs[i % 8], inverse=inv))
    print("%s %d limb-transforms: L2-resident %.1f us (%.3f us/limb)   HBM-streaming %.1f us (%.3f us/limb)" %
          ("intt" if inv else "ntt ", n, t_l2, t_l2 / n, t_hbm, t_hbm / n))
