"""torchrun worker: peer-direct limb-sharded key switch, one rank per GPU, NO collective on the data path — the base
conversions read the other GPUs' buffers over NVLink (cudaIpc mappings), ordered by epoch flags in peer memory.  One C-ABI
call per key switch (hml_keyswitch_sharded).  Checked against the unsharded GPU path; timed against the NCCL variant and the
single-GPU key switch.
Run as: python -m torch.distributed.run --nproc-per-node G --master-addr 127.0.0.1 tests/mp_sharded_keyswitch_p2p.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import homulator_b200 as hml  # noqa: E402
from orc import Oracle, uniform_limbs  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, ML, L, A = 65536, 45, 35, 15
    ctx = hml.Context(N=N, max_level=ML, alpha=A, device=local)
    o = Oracle(N, 36, ML, A)
    d = uniform_limbs(o.moduli[:L], N, 3000)
    evk = uniform_limbs(o.moduli[:L] + o.moduli[ML:], N, 3001, lead=(3, 2))

    def dev(x):
        return torch.from_numpy(np.ascontiguousarray(x).view(np.int64)).cuda()

    def exchange(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    sh = hml.Shard.ipc(ctx, L, rank, world, exchange)
    sh.prepare(L)
    d_own, evk_own = dev(d[sh.own_q(L)]), dev(evk[:, :, sh.own_ext(L)])
    dist.barrier()
    ref0, ref1 = ctx.keyswitch(L, dev(d), dev(evk))
    idx = torch.tensor(sh.own_q(L), device="cuda")
    ok = True
    for _ in range(3):  # buffer reuse across epochs
        o0, o1 = sh.keyswitch(L, d_own, evk_own)
        torch.cuda.synchronize()
        ok = ok and torch.equal(o0, ref0[idx]) and torch.equal(o1, ref1[idx])
    sh.check()

    def timeit(fn, reps=20):
        for _ in range(3):
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

    def all_gather(buf):
        dist.all_gather_into_tensor(buf, buf[rank].clone())

    o0b, o1b = ctx.empty(len(sh.own_q(L)), N), ctx.empty(len(sh.own_q(L)), N)
    t_p2p = timeit(lambda: sh.keyswitch(L, d_own, evk_own, o0b, o1b))
    t_nccl = timeit(lambda: ctx.keyswitch_sharded(L, d_own, evk_own, rank, world, all_gather))
    dd, ee = dev(d), dev(evk)
    t_one = timeit(lambda: ctx.keyswitch(L, dd, ee))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("sharded keyswitch world=%d: peer-direct %.1f us, NCCL all-gather %.1f us, single GPU %.1f us" % (world, t_p2p, t_nccl, t_one))
        print("SHARDED_P2P_OK" if int(flag) == 1 else "SHARDED_P2P_MISMATCH")
    dist.barrier()
    torch.cuda.synchronize()
    sh.close()
    dist.barrier()
    dist.destroy_process_group()
    return 0 if int(flag) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
