"""Profiling driver: a few single hmult / hrotate at the north-star config, nothing else on the GPU.

    python profiles/prof_hmult.py [hmult|hrotate|ntt] [n_warm] [n_prof]

Used under ncu (see profiles/README.md); kernel names all live in namespace hml::.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import homulator_b200 as hml  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "hmult"
n_warm = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_prof = int(sys.argv[3]) if len(sys.argv) > 3 else 2
L, A = 35, 15
ctx = hml.Context(os.path.join(ROOT, "config", "config_4.cfg"), 45, A)
q = list(range(L))
evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
a = ctx.uniform(q, 1, lead=(2,))
b = ctx.uniform(q, 2, lead=(2,))
torch.cuda.synchronize()
def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


if op == "time":
    idx = [ctx.ext_mod_idx(L)[i % 50] for i in range(115)]
    xs = [ctx.uniform(idx, 5 + i) for i in range(4)]
    y = ctx.empty(115, ctx.N)
    k = [0]

    def f_ntt():
        k[0] += 1
        ctx.ntt(xs[k[0] % 4], idx, out=y)

    def f_intt():
        k[0] += 1
        ctx.intt(xs[k[0] % 4], idx, out=y)

    t_ntt, t_intt = timeit(f_ntt), timeit(f_intt)
    print("ntt  115 limbs: %.1f us = %.3f us/limb = %.2f Mlimb/s" % (t_ntt, t_ntt / 115, 115 / t_ntt))
    print("intt 115 limbs: %.1f us = %.3f us/limb = %.2f Mlimb/s" % (t_intt, t_intt / 115, 115 / t_intt))
    print("hmult   %.1f us" % timeit(lambda: ctx.hmult(L, a, b, evk)))
    print("hrotate %.1f us" % timeit(lambda: ctx.hrotate(L, a, evk, 5)))
    d = ctx.uniform(q, 9)
    print("keyswitch %.1f us" % timeit(lambda: ctx.keyswitch(L, d, evk)))
elif op == "ntt":  # the launch shape bench.py's roofline times: 32 ciphertexts x 115 limbs per launch pair
    idx = [ctx.ext_mod_idx(L)[i % 50] for i in range(115)]
    x = ctx.uniform(idx, 5, lead=(32,))
    y = ctx.empty(32, 115, ctx.N)
    for _ in range(n_warm + n_prof):
        ctx.ntt_batch(x, idx, out=y)
        ctx.ntt_batch(x, idx, out=y, inverse=True)
else:
    for _ in range(n_warm + n_prof):
        if op == "hmult":
            ctx.hmult(L, a, b, evk)
        else:
            ctx.hrotate(L, a, evk, 5)
torch.cuda.synchronize()
print("done", op)
