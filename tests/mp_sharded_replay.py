"""torchrun worker: BASELINE.json configs[4] — the synthetic rotation-heavy op sequences (homulator_b200/replay.py) replayed
on LIMB-SHARDED operands, one rank per GPU, through the C ABI only (hml_replay_* over an hml_shard): every key switch and
rescale exchanges over NVLink without a collective; plain launches and ONE CUDA graph per rank.  Checked against the
single-GPU replay of the same trace; timed against it.
Run as: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tests/mp_sharded_replay.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import homulator_b200 as hml  # noqa: E402
from homulator_b200.replay import bsgs_trace, shard_operands, trace_counts  # noqa: E402


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

    def exchange(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    sh = hml.Shard.ipc(ctx, L, rank, world, exchange)
    sh.prepare(L)
    x_own, pts_own, keys_own, evk_own = shard_operands(sh, L, x, pts, keys, evk)
    dist.barrier()
    one = hml.Replay(ctx, L, tr).bind(x, pts, keys, evk)
    ref = one.run().result("z").clone()
    keep = torch.tensor([i for i in sh.own_q(L) if i < L - 1], device="cuda", dtype=torch.long)

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

    ok = True
    times = {}
    for name, graph in (("launches", False), ("graph", True)):
        rp = hml.Replay(ctx, L, tr, shard=sh, graph=graph).bind(x_own, pts_own, keys_own, evk_own)
        got = rp.run().result("z")
        torch.cuda.synchronize()
        ok = ok and torch.equal(got.contiguous(), ref[:, keep])
        times[name] = timeit(rp.run)
        got = rp.result("z")
        torch.cuda.synchronize()
        ok = ok and torch.equal(got.contiguous(), ref[:, keep])
        dist.barrier()
        rp.close()
    sh.check()
    t_one = timeit(one.run)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("op sequence %s world=%d: limb-sharded peer-direct %.1f us (as one CUDA graph per rank %.1f us), one GPU %.1f us"
              % (trace_counts(tr), world, times["launches"], times["graph"], t_one))
        print("SHARDED_REPLAY_OK" if int(flag) == 1 else "SHARDED_REPLAY_MISMATCH")
    dist.barrier()
    torch.cuda.synchronize()
    one.close()
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(flag) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
