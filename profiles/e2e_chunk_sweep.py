"""Host-buffer path: microseconds per hmult end to end (pinned host buffers, 32 ciphertexts per call) for the current
HML_HOST_CHUNK, uint64 words and the packed 5-byte format.  Run once per setting."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import homulator_b200 as hml
ctx = hml.Context(os.path.join(ROOT, "config", "config_4.cfg"), 45, 15)
L, n, N = 35, 32, 65536
q = list(range(L))
evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
ah = ctx.uniform(q, 1, lead=(n, 2)).cpu().pin_memory()
bh = ctx.uniform(q, 2, lead=(n, 2)).cpu().pin_memory()
oh = torch.empty(n, 2, L - 1, N, dtype=torch.int64).pin_memory()
ap, bp = ctx.pack_host(ah), ctx.pack_host(bh)
op = torch.empty(n * 2 * (L - 1) * 5 * N, dtype=torch.uint8).pin_memory()
def t(fn, reps=4):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps / n * 1e6
print("HML_HOST_CHUNK=%s words %.1f us  packed %.1f us per hmult" % (os.environ.get("HML_HOST_CHUNK", "default"),
      t(lambda: ctx.hmult_host(L, ah, bh, evk, oh)), t(lambda: ctx.hmult_host_packed(L, n, ap, bp, evk, op))))
