import os, sys, time, ctypes
sys.path.insert(0, os.getcwd())
import torch, homulator_b200 as hml
ctx = hml.Context(os.path.join(os.getcwd(), "config", "config_4.cfg"), 45, 15)
L, n = 35, 8
q = list(range(L))
evk = ctx.uniform(ctx.ext_mod_idx(L), 3, lead=(3, 2))
a = ctx.uniform(q, 1, lead=(n, 2)); b = ctx.uniform(q, 2, lead=(n, 2))
ah, bh = a.cpu().pin_memory(), b.cpu().pin_memory()
oh = torch.empty(n, 2, L - 1, 65536, dtype=torch.int64).pin_memory()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
# raw copies
d = torch.empty_like(a)
h2d = t(lambda: d.copy_(ah, non_blocking=True))
print("H2D torch-pinned %.1f GB/s" % (ah.numel() * 8 / h2d / 1e9))
o = ctx.empty(n, 2, L - 1, 65536)
d2h = t(lambda: oh.copy_(o, non_blocking=True))
print("D2H torch-pinned %.1f GB/s" % (oh.numel() * 8 / d2h / 1e9))
e = t(lambda: ctx.hmult_host(L, ah, bh, evk, oh))
print("e2e torch-pinned: %.1f us/hmult" % (e * 1e6 / n))
# cudaMallocHost buffers
lib = ctx.lib
def host_alloc(words):
    p = ctypes.c_void_p()
    assert lib.hml_host_alloc_pinned(ctx.h, ctypes.c_uint64(words), ctypes.byref(p)) == 0
    return p
lib.hml_host_alloc_pinned.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_void_p)]
na, no = ah.numel(), oh.numel()
pa, pb, po = host_alloc(na), host_alloc(na), host_alloc(no)
ctypes.memmove(pa, ah.data_ptr(), na * 8); ctypes.memmove(pb, bh.data_ptr(), na * 8)
lib.hml_hmult_host.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
e2 = t(lambda: lib.hml_hmult_host(ctx.h, L, n, pa, pb, ctypes.c_void_p(evk.data_ptr()), L, po))
print("e2e cudaMallocHost: %.1f us/hmult" % (e2 * 1e6 / n))
