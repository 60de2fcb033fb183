import os, sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, os.getcwd())
import torch, homulator_b200 as hml
L, A = 35, 15
ctx = hml.Context(os.path.join(os.getcwd(), "config", "config_4.cfg"), 45, A)
q = list(range(L))
evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
B = int(os.environ.get("B", "32"))
a = ctx.uniform(q, 1, lead=(B, 2)); b = ctx.uniform(q, 2, lead=(B, 2))
out = ctx.empty(B, 2, L - 1, ctx.N)
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
t = timeit(lambda: ctx.hmult_batch(L, a, b, evk, out=out))
print("hmult_batch B=%d: %.1f us/hmult" % (B, t / B))
o2 = ctx.empty(B, 2, L, ctx.N)
t = timeit(lambda: ctx.hrotate_batch(L, a, evk, 5, out=o2))
print("hrotate_batch B=%d: %.1f us/hrotate" % (B, t / B))
t = timeit(lambda: ctx.hmult(L, a[0], b[0], evk))
print("hmult single: %.1f us" % t)
