"""torchrun worker: BASELINE.json configs[4] — the synthetic baby-step/giant-step op sequence (homulator_b200/replay.py)
replayed on LIMB-SHARDED operands, one rank per GPU, every key switch and the rescale exchanging over NVLink without a
collective.  Checked against the single-GPU replay of the same trace; timed against it.
Run as: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tests/mp_sharded_replay.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import homulator_b200 as hml  # noqa: E402
from homulator_b200.replay import bsgs_trace, replay, replay_sharded, trace_counts  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, ML, L, A = 65536, 45, 35, 15
    ctx = hml.Context(N=N, max_level=ML, alpha=A, device=local)
    q = list(range(L))
    # the same seeded operands on every rank (full copies for the single-GPU replay, slices for the sharded one)
    x = ctx.uniform(q, 1, lead=(2,))
    evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
    tr = bsgs_trace(4, 4)
    rots = sorted({op[3] for op in tr if op[0] == "hrotate"})
    keys = {r: ctx.uniform(ctx.ext_mod_idx(L), 100 + r, lead=(3, 2)) for r in rots}
    pts = {i: ctx.uniform(q, 200 + i) for i in range(16)}
    lay = hml.shard_layout(L, A, rank, world)
    own = lay["own_q"]
    own_e = own + [L + j for j in lay["own_p"]]
    oi, oe = torch.tensor(own, device="cuda"), torch.tensor(own_e, device="cuda")
    x_own = x[:, oi].contiguous()
    evk_own = evk[:, :, oe].contiguous()
    keys_own = {r: k[:, :, oe].contiguous() for r, k in keys.items()}
    pts2 = {i: torch.stack([p[oi], p[oi]]).contiguous() for i, p in pts.items()}

    def exchange(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    sh = ctx.shard_p2p_setup(L, rank, world, exchange)
    dist.barrier()
    ref = replay(ctx, L, tr, x, pts, keys, evk)["z"]
    got = replay_sharded(sh, tr, x_own, pts2, keys_own, evk_own)["z"]
    torch.cuda.synchronize()
    keep = torch.tensor([i for i in own if i < L - 1], device="cuda")
    ok = torch.equal(got.contiguous(), ref[:, keep])

    def timeit(fn, reps=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    t_sh = timeit(lambda: replay_sharded(sh, tr, x_own, pts2, keys_own, evk_own))
    t_one = timeit(lambda: replay(ctx, L, tr, x, pts, keys, evk))
    # the same sharded sequence as ONE CUDA graph per rank: epochs come from device-side counters, so the graph replays;
    # the host enqueue cost (about 100 us per key switch through Python) no longer bounds the ranks' small kernels
    t_graph, ok_graph = float("nan"), True
    if os.environ.get("HML_SHARDED_GRAPH", "1") != "0":
        dist.barrier()
        torch.cuda.synchronize()
        sh2 = ctx.shard_p2p_setup(L, rank, world, exchange)
        sh2.device_epochs = True
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            for _ in range(2):
                z = replay_sharded(sh2, tr, x_own, pts2, keys_own, evk_own)["z"]
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            zg = replay_sharded(sh2, tr, x_own, pts2, keys_own, evk_own)["z"]
        torch.cuda.synchronize()
        dist.barrier()
        g.replay()
        torch.cuda.synchronize()
        ok_graph = torch.equal(zg.contiguous(), ref[:, keep])
        t_graph = timeit(lambda: g.replay())
        ok = ok and ok_graph
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("op sequence %s world=%d: limb-sharded peer-direct %.1f us (as one CUDA graph per rank %.1f us), one GPU %.1f us"
              % (trace_counts(tr), world, t_sh, t_graph, t_one))
        print("SHARDED_REPLAY_OK" if int(flag) == 1 else "SHARDED_REPLAY_MISMATCH")
    dist.barrier()
    torch.cuda.synchronize()
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(flag) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
