"""Profiling driver for the batched path: hmult_batch (or hrotate_batch) over one chunk of B (default 16) ciphertexts.

    python profiles/prof_batch.py [n_warm] [n_prof] [B] [hmult|hrotate]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import homulator_b200 as hml  # noqa: E402

n_warm = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n_prof = int(sys.argv[2]) if len(sys.argv) > 2 else 1
L, A = 35, 15
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
ctx = hml.Context(os.path.join(ROOT, "config", "config_4.cfg"), 45, A)
q = list(range(L))
evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
a = ctx.uniform(q, 1, lead=(B, 2))
b = ctx.uniform(q, 2, lead=(B, 2))
out = ctx.empty(B, 2, L - 1, ctx.N)
op = sys.argv[4] if len(sys.argv) > 4 else "hmult"
out_r = ctx.empty(B, 2, L, ctx.N) if op == "hrotate" else None
torch.cuda.synchronize()
for _ in range(n_warm + n_prof):
    if op == "hrotate":
        ctx.hrotate_batch(L, a, evk, 5, out=out_r)
    else:
        ctx.hmult_batch(L, a, b, evk, out=out)
torch.cuda.synchronize()
print("done")
