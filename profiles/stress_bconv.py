"""Repeatability stress of the tcgen05 base conversion and the batched ops: every repetition must be bit-identical to the
first (a race between the packer, the tensor core and the epilogue would show up as a flipped word)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import homulator_b200 as hml  # noqa: E402

ctx = hml.Context(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "config", "config_4.cfg"), 45, 15)
L = 35
bad = 0
for src, dst, nb in ((list(range(45, 60)), list(range(35)), 64), (list(range(15)), list(range(15, 35)) + list(range(45, 60)), 32),
                     (list(range(30, 35)), list(range(30)) + list(range(45, 60)), 32)):
    x = ctx.uniform(src, 5, lead=(nb,))
    ref = ctx.bconv_batch(x, src, dst)
    out = torch.empty_like(ref)
    for i in range(150):
        ctx.bconv_batch(x, src, dst, out=out)
        if not torch.equal(out, ref):
            bad += 1
            print("bconv %d->%d rep %d: %d words differ" % (len(src), len(dst), i, int((out != ref).sum())))
            break
    print("bconv %d -> %d x %d: 150 repetitions identical" % (len(src), len(dst), nb) if not bad else "MISMATCH", flush=True)
q = list(range(L))
evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
a, b = ctx.uniform(q, 1, lead=(32, 2)), ctx.uniform(q, 2, lead=(32, 2))
ref = ctx.hmult_batch(L, a, b, evk)
rot = ctx.hrotate_batch(L, a, evk, 5)
o1, o2 = torch.empty_like(ref), torch.empty_like(rot)
for i in range(40):
    ctx.hmult_batch(L, a, b, evk, out=o1)
    ctx.hrotate_batch(L, a, evk, 5, out=o2)
    if not (torch.equal(o1, ref) and torch.equal(o2, rot)):
        bad += 1
        print("op rep %d differs" % i)
        break
print("hmult_batch / hrotate_batch x 32: 40 repetitions identical" if not bad else "MISMATCH")
# one ciphertext per launch (tile queue of the column passes on few tiles per CTA, automorphism on load, fused inner product),
# interleaved with batched launches so that the queue counters change hands between launch shapes
r1, r2 = ctx.hmult(L, a[3], b[3], evk), ctx.hrotate(L, a[5], evk, 25)
p1, p2 = torch.empty_like(r1), torch.empty_like(r2)
idx = [ctx.ext_mod_idx(L)[i % 50] for i in range(115)]
xn = ctx.uniform(idx, 77, lead=(8,))
rn = ctx.ntt_batch(xn, idx)
pn = torch.empty_like(rn)
for i in range(100):
    ctx.hmult(L, a[3], b[3], evk, out=p1)
    ctx.ntt_batch(xn, idx, out=pn)
    ctx.hrotate(L, a[5], evk, 25, out=p2)
    if not (torch.equal(p1, r1) and torch.equal(p2, r2) and torch.equal(pn, rn)):
        bad += 1
        print("single-op rep %d differs" % i)
        break
print("hmult / hrotate / ntt_batch interleaved: 100 repetitions identical" if not bad else "MISMATCH")
sys.exit(1 if bad else 0)
