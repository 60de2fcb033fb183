"""CPU check of the NTT kernels' data movement (homulator_b200/csrc/ntt.cu) through the thread-level emulator in
tests/emu_ntt.py: thread maps, swizzled shared-memory tiles, cp.async chunk maps and the permuted row-twiddle blob must
reproduce the oracle's transform exactly for every two-pass ring size, in both directions and for both column-pass CTA
shapes, and the row pass's shared-memory accesses must be bank-conflict free."""
import numpy as np
import pytest

import emu_ntt
from orc import Oracle


@pytest.mark.parametrize("logN,NT", [(13, 256), (13, 128), (14, 128), (15, 256), (16, 128), (16, 256)])
def test_emulated_kernel_layout_matches_oracle(logN, NT):
    N = 1 << logN
    o = Oracle(N, 36, 2, 1)
    q, psi = o.moduli[1], o.psi[1]
    x = np.random.default_rng(logN).integers(0, q, N, dtype=np.uint64)
    got = emu_ntt.forward([int(v) for v in x], psi, q, logN, NT=NT)
    assert got == [int(v) for v in o.ntt(1, x)]
    assert emu_ntt.inverse(got, psi, q, logN, NT=NT) == [int(v) for v in x]


def test_bias_identity():
    """a constant added to coefficient 0 shows up unchanged in every evaluation slot (the canonicalisation trick of the
    forward column pass)"""
    N = 1 << 13
    o = Oracle(N, 36, 2, 1)
    q, psi = o.moduli[0], o.psi[0]
    x = np.random.default_rng(5).integers(0, q, N, dtype=np.uint64)
    h = (q - 1) // 2
    y = x.copy()
    y[0] = (int(y[0]) - h) % q
    a, b = o.ntt(0, x), o.ntt(0, y)
    assert np.array_equal((b.astype(object) + h) % q, a.astype(object))
