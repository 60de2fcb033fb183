"""Child process of tests/test_gpu_sharded.py::test_fused_exchange_with_device_side_epochs_on_two_streams (its own CUDA
context: a spin-wait that times out traps, which must not poison the rest of the suite)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import homulator_b200 as hml  # noqa: E402
from gpu_common import to_dev, to_host  # noqa: E402
from orc import Oracle, uniform_limbs  # noqa: E402


def main():
    N, ML, L, A, world = 8192, 9, 8, 3, 2
    o = Oracle(N, 36, ML, A)
    Oracle.set_threads(0)
    beta = -(-L // A)
    a = uniform_limbs(o.moduli[:L], N, 2300, lead=(2,))
    b = uniform_limbs(o.moduli[:L], N, 2301, lead=(2,))
    evk = uniform_limbs(o.moduli[:L] + o.moduli[ML:], N, 2302, lead=(beta, 2))
    want0, want1 = o.keyswitch(L, a[1], evk, L)
    want_mul = o.hmult(L, a, b, evk, L)
    Oracle.set_threads(1)
    ctxs = [hml.Context(N=N, max_level=ML, alpha=A) for _ in range(world)]
    lays = [hml.shard_layout(L, A, r, world) for r in range(world)]
    g1 = [c.dev_alloc(world * lays[r]["gather1_slots"] * N) for r, c in enumerate(ctxs)]
    g2 = [c.dev_alloc(world * 2 * lays[r]["gather2_slots"] * N) for r, c in enumerate(ctxs)]
    fl = [c.dev_alloc(3 * world + 8) for c in ctxs]
    rb = [c.dev_alloc(2 * N) for c in ctxs]
    shs = [hml.ShardP2P(c, L, r, world, g1, g2, fl, rb) for r, c in enumerate(ctxs)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    own = [lays[r]["own_q"] for r in range(world)]
    a_own = [to_dev(a[:, own[r]]) for r in range(world)]
    b_own = [to_dev(b[:, own[r]]) for r in range(world)]
    evk_own = [to_dev(evk[:, :, own[r] + [L + j for j in lays[r]["own_p"]]]) for r in range(world)]
    # warm-up in phase mode: builds the plans, offset tables and workspaces (their allocations synchronise the device, which
    # must not happen while a rank is spinning for a peer that the host has not enqueued yet), then the flags start over
    R = range(world)
    for r in R:
        shs[r].begin(a_own[r][1])
    for r in R:
        shs[r].mid(a_own[r][1], evk_own[r])
    cs = []
    for r in R:
        k0, k1 = shs[r].end()
        cs.append(shs[r].hmult_post(a_own[r][0], a_own[r][1], k0, k1))
    for r in R:
        shs[r].rescale_begin(cs[r])
    for r in R:
        shs[r].rescale_end(cs[r])
    torch.cuda.synchronize()
    for r in R:
        shs[r].reset_flags()
    ks, mul = [None] * world, [None] * world
    for r in range(world):
        shs[r].device_epochs = True
        with torch.cuda.stream(streams[r]):
            for _ in range(3):
                ks[r] = shs[r].keyswitch(a_own[r][1], evk_own[r])
            mul[r] = shs[r].hmult(a_own[r], b_own[r], evk_own[r])
    torch.cuda.synchronize()
    got0, got1 = np.zeros((L, N), dtype=np.uint64), np.zeros((L, N), dtype=np.uint64)
    gotm = np.zeros((2, L - 1, N), dtype=np.uint64)
    for r in range(world):
        got0[own[r]], got1[own[r]] = to_host(ks[r][0]), to_host(ks[r][1])
        gotm[:, [i for i in own[r] if i < L - 1]] = to_host(mul[r].contiguous())
    assert np.array_equal(got0, want0) and np.array_equal(got1, want1)
    assert np.array_equal(gotm, want_mul)
    print("FUSED_EXCHANGE_OK")


if __name__ == "__main__":
    main()
