"""Randomised parity sweep of hmult / hrotate / keyswitch / rescale against the scalar oracle over many (N, maxLevel, L, alpha)
shapes (digit counts 1..8, ragged last digits, L < alpha, 1..48 conversion targets): python profiles/fuzz_parity.py [n_cases] [seed] [maxML]
(maxML > 14 keeps the rings small and reaches the 2- and 3-slab conversion kernels: up to 48 sources)"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import homulator_b200 as hml  # noqa: E402
from gpu_common import to_dev, to_host  # noqa: E402
from orc import Oracle, uniform_limbs  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rnd = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
ml_max = int(sys.argv[3]) if len(sys.argv) > 3 else 14
bad = 0
Oracle.set_threads(0)
for case in range(n_cases):
    N = rnd.choice([128, 256, 512, 2048, 8192, 8192, 16384] if ml_max <= 14 else [128, 256, 512, 8192])
    ML = rnd.randint(2, ml_max)
    A = rnd.randint(1, ML)
    L = rnd.randint(2, ML)
    if -(-L // A) > 8:
        A = -(-L // 8)
    beta = -(-L // A)
    ctx, o = hml.Context(N=N, max_level=ML, alpha=A), Oracle(N, 36, ML, A)
    a = uniform_limbs(o.moduli[:L], N, 10 * case + 1, lead=(2,))
    b = uniform_limbs(o.moduli[:L], N, 10 * case + 2, lead=(2,))
    evk = uniform_limbs(o.moduli[:L] + o.moduli[ML:], N, 10 * case + 3, lead=(beta, 2))
    g = pow(5, rnd.randint(1, 40), 2 * N)
    ok = np.array_equal(to_host(ctx.hmult(L, to_dev(a), to_dev(b), to_dev(evk))), o.hmult(L, a, b, evk, L))
    ok &= np.array_equal(to_host(ctx.hrotate(L, to_dev(a), to_dev(evk), g)), o.hrotate(L, a, evk, L, g))
    nb = rnd.randint(2, 5)
    As = torch.stack([to_dev(a if i % 2 == 0 else b) for i in range(nb)])
    got = ctx.hmult_batch(L, As, As.flip(0).contiguous(), to_dev(evk))
    i = rnd.randrange(nb)
    ok &= np.array_equal(to_host(got[i]), o.hmult(L, to_host(As[i]), to_host(As[nb - 1 - i]), evk, L))
    print("case %2d N=%5d maxL=%2d L=%2d alpha=%2d beta=%d: %s" % (case, N, ML, L, A, beta, "ok" if ok else "MISMATCH"), flush=True)
    bad += 0 if ok else 1
    ctx.close()
print("FUZZ_OK" if not bad else "FUZZ_MISMATCH %d" % bad)
sys.exit(1 if bad else 0)
