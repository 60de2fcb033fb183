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


def _brev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


@pytest.mark.parametrize("logN,g", [(13, 5), (13, 2 * 8192 - 1), (14, 25), (15, 3125), (16, 5), (16, 25), (16, 2 * 65536 - 1), (16, 3 ** 9)])
def test_automorphism_maps_rows_onto_rows(logN, g):
    """The row structure hrotate's fused automorphism relies on (homulator_b200/csrc/ntt_core.cuh RowSigma / sigma_src /
    sigma_dst): in evaluation order out[k] = in[k'] with every 256-slot row R reading ONE source row R', the slot inside it
    an affine permutation of brev8(klo).  The integer formulas of the kernels, restated here, must reproduce the oracle's
    index map (orc_automorph_index) for every slot."""
    N, h = 1 << logN, logN - 8
    o = Oracle(N, 36, 2, 1)
    want = o.automorph_index(g)  # out[k] = in[want[k]]
    g &= 2 * N - 1
    ginv8 = pow(g, -1, 256)
    for R in range(N >> 8):
        b = _brev(R, h)
        t = (g * b + ((g - 1) >> 1)) & 0xFFFFFFFF        # 32-bit arithmetic, as on the device
        src_row, d = _brev(t & ((1 << h) - 1), h), t >> h
        for klo in range(256):
            kp = _brev((g * _brev(klo, 8) + d) & 255, 8)   # sigma_src
            assert want[R * 256 + klo] == src_row * 256 + kp
            assert _brev((ginv8 * (_brev(kp, 8) - d)) & 255, 8) == klo   # sigma_dst inverts it


def test_column_tile_queue_enumeration():
    """The column passes' tile queue (ntt.cu col_decode) numbers the work items of a launch without the skipped (digit-owned)
    polys: tile fastest, then the members of a ciphertext — n_skip leading limbs with n_polys - 1 members, the others with
    n_polys — then the ciphertexts.  Restated here and compared with a brute-force walk of the (batch, poly, limb, tile) lattice."""
    for tiles, n_limbs, n_polys, n_batch, n_skip in ((16, 50, 3, 2, 35), (32, 7, 2, 3, 7), (16, 5, 1, 4, 0), (16, 10, 4, 1, 3)):
        skip = [(e * 7) % n_polys if e < n_skip else 0xFF for e in range(n_limbs)]
        pairs = n_skip * (n_polys - 1) + (n_limbs - n_skip) * n_polys
        total = pairs * n_batch * tiles
        got = set()
        for wi in range(total):
            tile, r = wi % tiles, wi // tiles
            batch, rr = r // pairs, r % pairs
            own = n_skip * (n_polys - 1)
            if rr < own:
                limb, poly = rr // (n_polys - 1), rr % (n_polys - 1)
                poly += poly >= skip[limb]
            else:
                rr -= own
                limb, poly = n_skip + rr // n_polys, rr % n_polys
            assert poly != skip[limb]
            got.add((batch, poly, limb, tile))
        want = {(b, p, e, t) for b in range(n_batch) for p in range(n_polys) for e in range(n_limbs) for t in range(tiles) if p != skip[e]}
        assert got == want and len(got) == total
