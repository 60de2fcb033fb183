"""CPU checks of the tcgen05 base conversion's host side (homulator_b200/csrc/bconv_umma.cu): the int8 operand image, the K
ordering the device packs the sources in, and the exactness of the epilogue's 5-level shift-add + one FP64 reduction.
`hml_dbg_bconv_umma_model` runs the kernel's data path on the CPU (same image bytes, same addressing, same arithmetic); it is
compared with big-integer arithmetic.  The hardware-side layout (descriptors, TMEM mapping) is covered by the GPU tests."""
import ctypes as C
import random

import numpy as np
import pytest

from homulator_b200 import api as hml


def _model(hat, n_src, n_dst, dst_q, fold, fold_q, y):
    lib = C.CDLL(hml.lib_path())
    f = lib.hml_dbg_bconv_umma_model
    u64p = C.POINTER(C.c_uint64)
    f.argtypes = [u64p, C.c_int, C.c_int, u64p, u64p, C.c_uint64, u64p, C.c_int, u64p]
    M = y.shape[1]
    out = np.zeros((n_dst, M), dtype=np.uint64)
    hat = np.ascontiguousarray(hat, dtype=np.uint64)
    dq = np.ascontiguousarray(dst_q, dtype=np.uint64)
    fo = np.ascontiguousarray(fold, dtype=np.uint64) if fold is not None else None
    rc = f(hat.ctypes.data_as(u64p), n_src, n_dst, dq.ctypes.data_as(u64p), fo.ctypes.data_as(u64p) if fo is not None else None,
           fold_q, np.ascontiguousarray(y).ctypes.data_as(u64p), M, out.ctypes.data_as(u64p))
    return rc, out


@pytest.mark.parametrize("n_src,n_dst,fold,extreme", [
    (15, 35, False, False), (15, 45, False, True), (5, 45, False, False), (16, 34, True, False), (16, 34, True, True),
    (28, 28, False, True), (48, 48, False, True), (1, 1, False, False), (17, 3, True, True), (33, 8, False, False),
])
def test_model_matches_big_integers(n_src, n_dst, fold, extreme):
    rnd = random.Random(n_src * 1000 + n_dst + fold)
    top = (1 << 36) - 1
    dst_q = [top - 2 * rnd.randrange(1000) if extreme else (rnd.randrange(1 << 35, 1 << 36) | 1) for _ in range(n_dst)]
    src_q = [top - 2 * rnd.randrange(1000) for _ in range(n_src)]
    fold_q = top - 4 if fold else 0
    M = 24
    hat = [[(dst_q[t] - 1 - rnd.randrange(3)) if extreme else rnd.randrange(dst_q[t]) for t in range(n_dst)] for _ in range(n_src)]
    if fold:
        for t in range(n_dst):
            hat[n_src - 1][t] = 1
    fo = [(fold_q - 1 - rnd.randrange(3)) if extreme else rnd.randrange(fold_q) for _ in range(n_src - 1)] if fold else None
    y = np.zeros((n_src, M), dtype=np.uint64)
    for i in range(n_src):
        for m in range(M):
            y[i, m] = src_q[i] - 1 - rnd.randrange(2) if (extreme or m < 2) else rnd.randrange(src_q[i])
    if fold:
        y[n_src - 1] = [fold_q - 1 if m % 2 else rnd.randrange(fold_q) for m in range(M)]
    rc, got = _model(hat, n_src, n_dst, dst_q, fo, fold_q, y)
    assert rc == 0
    for m in range(M):
        ys = [int(v) for v in y[:, m]]
        if fold:
            r = (ys[-1] + sum(a * b for a, b in zip(ys[:-1], fo))) % fold_q
            want = [(sum(ys[i] * hat[i][t] for i in range(n_src - 1)) + r) % dst_q[t] for t in range(n_dst)]
        else:
            want = [sum(ys[i] * hat[i][t] for i in range(n_src)) % dst_q[t] for t in range(n_dst)]
        assert [int(v) for v in got[:, m]] == want


def test_shapes_outside_the_tensor_core_path_are_refused():
    # 49 targets need 5 * 56 = 280 accumulator columns (> 256); 49 sources would let the 5-level sum pass 2^52
    for n_src, n_dst in ((15, 49), (49, 8)):
        rc, _ = _model([[1] * n_dst] * n_src, n_src, n_dst, [(1 << 36) - 5] * n_dst, None, 0, np.ones((n_src, 4), dtype=np.uint64))
        assert rc == 1
