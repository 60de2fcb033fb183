import os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, homulator_b200 as hml
N, ML, L, A = 65536, 45, 35, 15
ctx = hml.Context(N=N, max_level=ML, alpha=A)
d = ctx.uniform(list(range(L)), 1)
evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
sh = ctx.shard_p2p_setup(L, 0, 1, lambda o: [o])
for _ in range(3): sh.keyswitch(d, evk)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): sh.keyswitch(d, evk)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("world=1 sharded KS: cpu enqueue %.1f us per key switch, total %.1f us" % ((t1 - t0) / 50 * 1e6, (t2 - t0) / 50 * 1e6))
t0 = time.perf_counter()
for _ in range(50): ctx.keyswitch(L, d, evk)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("unsharded KS: cpu enqueue %.1f us" % ((t1 - t0) / 50 * 1e6))
# small-kernel latency floor: the same with tiny work (L=2 ring) -> per-launch GPU latency
