"""Randomised parity sweep (the round-1 profiles/fuzz_parity.py, now collected by pytest and actually covering keyswitch and
rescale): hmult / hrotate / keyswitch / rescale / batched hmult against the scalar oracle over seeded random
(N, maxLevel, L, alpha) shapes — digit counts 1..8, ragged last digits, L < alpha, 1..48 conversion sources and targets."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import homulator_b200 as hml  # noqa: E402
from gpu_common import to_dev, to_host  # noqa: E402
from orc import Oracle, uniform_limbs  # noqa: E402


def _shape(rnd, ml_max, rings):
    N = rnd.choice(rings)
    ML = rnd.randint(2, ml_max)
    A = rnd.randint(1, ML)
    L = rnd.randint(2, ML)
    if -(-L // A) > 8:
        A = -(-L // 8)
    return N, ML, A, L


@pytest.mark.parametrize("seed,ml_max,rings,n_cases", [
    (1234, 14, [128, 256, 512, 2048, 8192, 8192, 16384], 14),
    (99, 40, [128, 256, 512, 8192], 8),       # up to 48 conversion sources: the 2- and 3-slab tcgen05 kernels
])
def test_fuzz_parity(seed, ml_max, rings, n_cases):
    rnd = random.Random(seed)
    Oracle.set_threads(0)
    bad = []
    for case in range(n_cases):
        N, ML, A, L = _shape(rnd, ml_max, rings)
        beta = -(-L // A)
        ctx, o = hml.Context(N=N, max_level=ML, alpha=A), Oracle(N, 36, ML, A)
        a = uniform_limbs(o.moduli[:L], N, 10 * case + 1, lead=(2,))
        b = uniform_limbs(o.moduli[:L], N, 10 * case + 2, lead=(2,))
        evk = uniform_limbs(o.moduli[:L] + o.moduli[ML:], N, 10 * case + 3, lead=(beta, 2))
        g = pow(5, rnd.randint(1, 40), 2 * N)
        K = to_dev(evk)
        ok = np.array_equal(to_host(ctx.hmult(L, to_dev(a), to_dev(b), K)), o.hmult(L, a, b, evk, L))
        ok &= np.array_equal(to_host(ctx.hrotate(L, to_dev(a), K, g)), o.hrotate(L, a, evk, L, g))
        w0, w1 = o.keyswitch(L, b[1], evk, L)
        g0, g1 = ctx.keyswitch(L, to_dev(b[1]), K, L)
        ok &= np.array_equal(to_host(g0), w0) and np.array_equal(to_host(g1), w1)
        ok &= np.array_equal(to_host(ctx.rescale(L, to_dev(a[1]))), o.rescale(L, a[1]))
        nb = rnd.randint(2, 5)
        As = torch.stack([to_dev(a if i % 2 == 0 else b) for i in range(nb)])
        got = ctx.hmult_batch(L, As, As.flip(0).contiguous(), K)
        i = rnd.randrange(nb)
        ok &= np.array_equal(to_host(got[i]), o.hmult(L, to_host(As[i]), to_host(As[nb - 1 - i]), evk, L))
        if not ok:
            bad.append((N, ML, L, A))
        ctx.close()
    Oracle.set_threads(1)
    assert not bad, bad
