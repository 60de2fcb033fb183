"""GPU parity of the limb-sharded key switch (SURVEY.md 8e mode 2).

On one GPU the ranks are emulated one after another (a context per rank, the all-gathers done as device copies), so the
sharded arithmetic is checked bit-exactly against the oracle on every box; with >= 2 GPUs the real NCCL path
(tests/mp_sharded_keyswitch.py under torch.distributed.run) is exercised as well."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import homulator_b200 as hml  # noqa: E402
from gpu_common import to_dev, to_host  # noqa: E402
from orc import Oracle, uniform_limbs  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sharded_keyswitch_emulated(ctxs, L, d, evk, world):
    """d [L][N], evk [beta][2][L+alpha][N] numpy; returns (out0, out1) [L][N] assembled from all ranks."""
    A, N = ctxs[0].alpha, ctxs[0].N
    lays = [hml.shard_layout(L, A, r, world) for r in range(world)]
    d_own, evk_own, g1, g2 = [], [], [], []
    for r, lay in enumerate(lays):
        own_e = lay["own_q"] + [L + j for j in lay["own_p"]]
        d_own.append(to_dev(d[lay["own_q"]] if lay["own_q"] else np.zeros((1, N), dtype=np.uint64)))
        evk_own.append(to_dev(evk[:, :, own_e]) if own_e else None)
        g1.append(torch.zeros(world, lay["gather1_slots"], N, dtype=torch.int64, device="cuda"))
        g2.append(torch.zeros(world, 2, lay["gather2_slots"], N, dtype=torch.int64, device="cuda"))
    st = torch.cuda.current_stream().cuda_stream
    for r, c in enumerate(ctxs):
        c._chk(c.lib.hml_keyswitch_shard_begin(c.h, L, r, world, d_own[r].data_ptr(), g1[r].data_ptr(), st))
    for r in range(world):  # all-gather 1
        for s in range(world):
            g1[r][s].copy_(g1[s][s])
    for r, c in enumerate(ctxs):
        if evk_own[r] is None:
            continue
        c._chk(c.lib.hml_keyswitch_shard_mid(c.h, L, r, world, d_own[r].data_ptr(), g1[r].data_ptr(), evk_own[r].data_ptr(),
                                             g2[r].data_ptr(), st))
    for r in range(world):  # all-gather 2
        for s in range(world):
            g2[r][s].copy_(g2[s][s])
    out0, out1 = np.zeros((L, N), dtype=np.uint64), np.zeros((L, N), dtype=np.uint64)
    for r, c in enumerate(ctxs):
        nq = len(lays[r]["own_q"])
        o0 = torch.zeros(max(nq, 1), N, dtype=torch.int64, device="cuda")
        o1 = torch.zeros_like(o0)
        c._chk(c.lib.hml_keyswitch_shard_end(c.h, L, r, world, g2[r].data_ptr(), o0.data_ptr(), o1.data_ptr(), st))
        if nq:
            out0[lays[r]["own_q"]] = to_host(o0)[:nq]
            out1[lays[r]["own_q"]] = to_host(o1)[:nq]
    return out0, out1


@pytest.mark.parametrize("N,ML,L,A,world", [(256, 7, 7, 3, 2), (256, 7, 7, 3, 3), (1024, 6, 5, 2, 4), (64, 5, 2, 3, 2),
                                             (4096, 4, 4, 4, 8), (8192, 9, 8, 3, 2), (65536, 45, 35, 15, 4)])
def test_sharded_keyswitch_emulated_ranks(N, ML, L, A, world):
    o = Oracle(N, 36, ML, A)
    Oracle.set_threads(0)
    beta = -(-L // A)
    d = uniform_limbs(o.moduli[:L], N, 2000 + L)
    evk = uniform_limbs(o.moduli[:L] + o.moduli[ML:], N, 2001, lead=(beta, 2))
    want0, want1 = o.keyswitch(L, d, evk, L)
    Oracle.set_threads(1)
    ctxs = [hml.Context(N=N, max_level=ML, alpha=A) for _ in range(world)]
    got0, got1 = sharded_keyswitch_emulated(ctxs, L, d, evk, world)
    assert np.array_equal(got0, want0) and np.array_equal(got1, want1)
    # world = 1 degenerates to the unsharded schedule
    one0, one1 = sharded_keyswitch_emulated(ctxs[:1], L, d, evk, 1)
    assert np.array_equal(one0, want0) and np.array_equal(one1, want1)


def sharded_keyswitch_p2p_emulated(ctxs, L, d, evk, world, reps=2):
    """Peer-direct variant with the ranks emulated on one GPU and one stream: every rank's conversions read the other ranks'
    gather buffers through per-source offsets, ordered by the epoch flags (signals are enqueued before the waits that need
    them).  Runs `reps` key switches back to back (buffer reuse across epochs)."""
    A, N = ctxs[0].alpha, ctxs[0].N
    lays = [hml.shard_layout(L, A, r, world) for r in range(world)]
    g1 = [c.dev_alloc(world * lays[r]["gather1_slots"] * N) for r, c in enumerate(ctxs)]
    g2 = [c.dev_alloc(world * 2 * lays[r]["gather2_slots"] * N) for r, c in enumerate(ctxs)]
    fl = [c.dev_alloc(2 * world) for c in ctxs]
    sh = [hml.ShardP2P(c, L, r, world, g1, g2, fl) for r, c in enumerate(ctxs)]
    d_own, evk_own = [], []
    for r, lay in enumerate(lays):
        own_e = lay["own_q"] + [L + j for j in lay["own_p"]]
        d_own.append(to_dev(d[lay["own_q"]] if lay["own_q"] else np.zeros((1, N), dtype=np.uint64)))
        evk_own.append(to_dev(evk[:, :, own_e]) if own_e else None)
    for _ in range(reps):
        for r in range(world):
            sh[r].begin(d_own[r])
        for r in range(world):
            sh[r].mid(d_own[r], evk_own[r])
        outs = [sh[r].end() for r in range(world)]
    out0, out1 = np.zeros((L, N), dtype=np.uint64), np.zeros((L, N), dtype=np.uint64)
    for r in range(world):
        if lays[r]["own_q"]:
            out0[lays[r]["own_q"]] = to_host(outs[r][0])
            out1[lays[r]["own_q"]] = to_host(outs[r][1])
    return out0, out1


@pytest.mark.parametrize("N,ML,L,A,world", [(256, 7, 7, 3, 2), (256, 7, 7, 3, 3), (1024, 6, 5, 2, 4), (8192, 9, 8, 3, 2),
                                             (65536, 45, 35, 15, 4)])
def test_sharded_keyswitch_peer_direct_emulated_ranks(N, ML, L, A, world):
    o = Oracle(N, 36, ML, A)
    Oracle.set_threads(0)
    beta = -(-L // A)
    d = uniform_limbs(o.moduli[:L], N, 2100 + L)
    evk = uniform_limbs(o.moduli[:L] + o.moduli[ML:], N, 2101, lead=(beta, 2))
    want0, want1 = o.keyswitch(L, d, evk, L)
    Oracle.set_threads(1)
    ctxs = [hml.Context(N=N, max_level=ML, alpha=A) for _ in range(world)]
    got0, got1 = sharded_keyswitch_p2p_emulated(ctxs, L, d, evk, world)
    assert np.array_equal(got0, want0) and np.array_equal(got1, want1)


def _emulated_shards(ctxs, L, world):
    N = ctxs[0].N
    lays = [hml.shard_layout(L, ctxs[0].alpha, r, world) for r in range(world)]
    g1 = [c.dev_alloc(world * lays[r]["gather1_slots"] * N) for r, c in enumerate(ctxs)]
    g2 = [c.dev_alloc(world * 2 * lays[r]["gather2_slots"] * N) for r, c in enumerate(ctxs)]
    fl = [c.dev_alloc(3 * world) for c in ctxs]
    rb = [c.dev_alloc(2 * N) for c in ctxs]
    return [hml.ShardP2P(c, L, r, world, g1, g2, fl, rb) for r, c in enumerate(ctxs)], lays


@pytest.mark.parametrize("N,ML,L,A,world", [(256, 7, 7, 3, 2), (1024, 6, 5, 2, 3), (8192, 9, 8, 3, 4), (8192, 9, 9, 3, 2)])
def test_sharded_hrotate_and_hmult_peer_direct_emulated_ranks(N, ML, L, A, world):
    """Whole ops on limb-sharded ciphertexts (automorphism / tensor product / additions limb-local, key switch and rescale with
    peer-direct exchanges), ranks emulated phase by phase on one stream, against the oracle's unsharded hrotate and hmult."""
    o = Oracle(N, 36, ML, A)
    Oracle.set_threads(0)
    beta = -(-L // A)
    a = uniform_limbs(o.moduli[:L], N, 2200, lead=(2,))
    b = uniform_limbs(o.moduli[:L], N, 2201, lead=(2,))
    evk = uniform_limbs(o.moduli[:L] + o.moduli[ML:], N, 2202, lead=(beta, 2))
    g = 25
    want_rot = o.hrotate(L, a, evk, L, g)
    want_mul = o.hmult(L, a, b, evk, L)
    Oracle.set_threads(1)
    ctxs = [hml.Context(N=N, max_level=ML, alpha=A) for _ in range(world)]
    shs, lays = _emulated_shards(ctxs, L, world)
    R = range(world)
    own = [lays[r]["own_q"] for r in R]
    a_own = [to_dev(a[:, own[r]]) for r in R]
    b_own = [to_dev(b[:, own[r]]) for r in R]
    evk_own = [to_dev(evk[:, :, own[r] + [L + j for j in lays[r]["own_p"]]]) for r in R]
    # hrotate, twice (epochs)
    for _ in range(2):
        sig = [shs[r].hrotate_pre(a_own[r], g) for r in R]
        outs = [ctxs[r].empty(2, shs[r].nq, N) for r in R]
        for r in R:
            shs[r].begin(sig[r][1])
        for r in R:
            shs[r].mid(sig[r][1], evk_own[r])
        for r in R:
            k0, _ = shs[r].end(o1=outs[r][1])
            shs[r].hrotate_post(sig[r], k0, outs[r])
    got = np.zeros((2, L, N), dtype=np.uint64)
    for r in R:
        got[:, own[r]] = to_host(outs[r])
    assert np.array_equal(got, want_rot)
    # hmult
    pre = [shs[r].hmult_pre(a_own[r], b_own[r]) for r in R]
    for r in R:
        shs[r].begin(pre[r][2])
    for r in R:
        shs[r].mid(pre[r][2], evk_own[r])
    cs = []
    for r in R:
        k0, k1 = shs[r].end()
        cs.append(shs[r].hmult_post(pre[r][0], pre[r][1], k0, k1))
    for r in R:
        shs[r].rescale_begin(cs[r])
    res = [shs[r].rescale_end(cs[r]) for r in R]
    got = np.zeros((2, L - 1, N), dtype=np.uint64)
    for r in R:
        keep = [i for i in own[r] if i < L - 1]
        if keep:
            got[:, keep] = to_host(res[r].contiguous())
    assert np.array_equal(got, want_mul)


def test_fused_exchange_with_device_side_epochs_on_two_streams():
    """The one-launch signal + wait (hml_shard_sync) with device-side epoch counters — the form the multi-GPU path and its CUDA
    graph use — on ONE GPU: two emulated ranks, each on its own stream, so that a rank's spin-wait runs while the other rank's
    kernels make progress.  Three key switches and one hmult back to back, against the oracle (tests/sp_fused_exchange.py,
    in a child process: a spin-wait that times out traps and would take this process's CUDA context with it)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sp_fused_exchange.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FUSED_EXCHANGE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_peer_direct_needs_the_tcgen05_conversion():
    ctx = hml.Context(N=64, max_level=5, alpha=3)  # N < 128: the FP64 tensor-core kernel has no per-source offsets
    g1, g2, fl = ctx.dev_alloc(2 * 1 * 64), ctx.dev_alloc(2 * 2 * 2 * 64), ctx.dev_alloc(4)
    sh = hml.ShardP2P(ctx, 2, 0, 2, [g1, g1], [g2, g2], [fl, fl])
    d = ctx.uniform([0], 1)
    evk = ctx.uniform([0, 1, 2], 2, lead=(1, 2))
    with pytest.raises(hml.HmlError):
        ctx._chk(ctx.lib.hml_keyswitch_shard_mid_p2p(ctx.h, 2, 0, 2, d.data_ptr(), sh.p1, evk.data_ptr(), g2, None))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_sharded_keyswitch_nccl():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(ROOT, "tests", "mp_sharded_keyswitch.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_KS_OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_sharded_keyswitch_peer_direct_over_nvlink():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mp_sharded_keyswitch_p2p.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_P2P_OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs (gpurun --gpus 2)")
def test_sharded_op_sequence_over_nvlink():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29539", os.path.join(ROOT, "tests", "mp_sharded_replay.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDED_REPLAY_OK" in r.stdout
