"""ctypes binding of the scalar C oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "liboracle.so")


def build():
    src = os.path.join(ROOT, "oracle", "oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True,
                       capture_output=True)
    return _SO


_lib = None
_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32] * 4
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_n_moduli.restype = C.c_uint32
        L.orc_n_moduli.argtypes = [C.c_void_p]
        L.orc_modulus.restype = C.c_uint64
        L.orc_modulus.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_psi.restype = C.c_uint64
        L.orc_psi.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_set_direct.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        for f in ("orc_ntt", "orc_intt"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_uint32, _u64p]
        for f in ("orc_ntt_direct", "orc_intt_direct"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p]
        L.orc_negacyclic_schoolbook.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, _u64p]
        L.orc_ewe.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, _u64p, _u64p, C.c_int, _u64p]
        L.orc_automorph_eval.argtypes = [C.c_void_p, C.c_uint64, _u64p, _u64p]
        L.orc_automorph_coeff.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, _u64p, _u64p]
        L.orc_automorph_index.argtypes = [C.c_void_p, C.c_uint64, _u32p]
        L.orc_bconv.argtypes = [C.c_void_p, _u32p, C.c_uint32, C.c_uint32, _u64p, _u64p]
        L.orc_keyswitch.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, C.c_uint32, _u64p, _u64p]
        L.orc_rescale.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p]
        L.orc_modup.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p]
        L.orc_keyswitch_digits.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, C.c_uint32, _u64p, _u64p]
        L.orc_hrotate_hoisted.argtypes = [C.c_void_p, C.c_uint32, _u64p, C.c_uint32, C.POINTER(_u64p), C.c_uint32, _u64p,
                                          C.POINTER(_u64p)]
        L.orc_hmult.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, _u64p, C.c_uint32, _u64p]
        L.orc_hrotate.argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, C.c_uint32, C.c_uint64, _u64p]
        for f in ("orc_hadd", "orc_pmult", "orc_padd"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_uint32, _u64p, _u64p, _u64p]
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


class Oracle:
    def __init__(self, N, word_bits, max_level, alpha):
        self.N, self.word_bits, self.max_level, self.alpha = N, word_bits, max_level, alpha
        self.h = lib().orc_create(N, word_bits, max_level, alpha)
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.n_mod = max_level + alpha
        self.moduli = [int(lib().orc_modulus(self.h, i)) for i in range(self.n_mod)]
        self.psi = [int(lib().orc_psi(self.h, i)) for i in range(self.n_mod)]

    def __del__(self):
        try:
            if self.h:
                lib().orc_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_direct(self, flag):
        lib().orc_set_direct(self.h, int(flag))

    @staticmethod
    def set_threads(n):
        return lib().orc_set_threads(n)

    def ext_mod_idx(self, L):
        return list(range(L)) + [self.max_level + j for j in range(self.alpha)]

    def ntt(self, mi, a):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().orc_ntt(self.h, mi, _p(a))
        return a

    def intt(self, mi, a):
        a = np.ascontiguousarray(a, dtype=np.uint64).copy()
        lib().orc_intt(self.h, mi, _p(a))
        return a

    def ntt_direct(self, mi, a):
        out = np.empty(self.N, dtype=np.uint64)
        lib().orc_ntt_direct(self.h, mi, _p(np.ascontiguousarray(a)), _p(out))
        return out

    def intt_direct(self, mi, a):
        out = np.empty(self.N, dtype=np.uint64)
        lib().orc_intt_direct(self.h, mi, _p(np.ascontiguousarray(a)), _p(out))
        return out

    def schoolbook(self, mi, a, b):
        out = np.empty(self.N, dtype=np.uint64)
        lib().orc_negacyclic_schoolbook(self.h, mi, _p(a), _p(b), _p(out))
        return out

    def ewe(self, mi, x1, x2, x3, x4, sub=False):
        out = np.empty(self.N, dtype=np.uint64)
        lib().orc_ewe(self.h, mi, _p(x1), _p(x2), _p(x3), _p(x4), int(sub), _p(out))
        return out

    def automorph_eval(self, g, a):
        out = np.empty(self.N, dtype=np.uint64)
        lib().orc_automorph_eval(self.h, g, _p(np.ascontiguousarray(a)), _p(out))
        return out

    def automorph_coeff(self, mi, g, a):
        out = np.empty(self.N, dtype=np.uint64)
        lib().orc_automorph_coeff(self.h, mi, g, _p(np.ascontiguousarray(a)), _p(out))
        return out

    def automorph_index(self, g):
        out = np.empty(self.N, dtype=np.uint32)
        lib().orc_automorph_index(self.h, g, out.ctypes.data_as(_u32p))
        return out

    def bconv(self, src_idx, dst_idx, x):
        src = np.asarray(src_idx, dtype=np.uint32)
        out = np.empty(self.N, dtype=np.uint64)
        lib().orc_bconv(self.h, src.ctypes.data_as(_u32p), len(src), dst_idx, _p(np.ascontiguousarray(x)), _p(out))
        return out

    def keyswitch(self, L, d, evk, evk_q_limbs):
        o0 = np.empty((L, self.N), dtype=np.uint64)
        o1 = np.empty((L, self.N), dtype=np.uint64)
        lib().orc_keyswitch(self.h, L, _p(d), _p(evk), evk_q_limbs, _p(o0), _p(o1))
        return o0, o1

    def rescale(self, L, x):
        out = np.empty((L - 1, self.N), dtype=np.uint64)
        lib().orc_rescale(self.h, L, _p(x), _p(out))
        return out

    def hmult(self, L, a, b, evk, evk_q_limbs):
        out = np.empty((2, L - 1, self.N), dtype=np.uint64)
        lib().orc_hmult(self.h, L, _p(a), _p(b), _p(evk), evk_q_limbs, _p(out))
        return out

    def hrotate(self, L, ct, rk, evk_q_limbs, g):
        out = np.empty((2, L, self.N), dtype=np.uint64)
        lib().orc_hrotate(self.h, L, _p(ct), _p(rk), evk_q_limbs, g, _p(out))
        return out

    def hrotate_hoisted(self, L, ct, rks, evk_q_limbs, gs):
        """rotations gs (galois elements) of ONE ciphertext sharing one ModUp; returns a list of [2][L][N] arrays"""
        n = len(gs)
        outs = [np.empty((2, L, self.N), dtype=np.uint64) for _ in range(n)]
        keys = (_u64p * n)(*[_p(k) for k in rks])
        po = (_u64p * n)(*[_p(o) for o in outs])
        g = np.asarray(gs, dtype=np.uint64)
        lib().orc_hrotate_hoisted(self.h, L, _p(ct), n, keys, evk_q_limbs, _p(g), po)
        return outs

    def modup(self, L, d):
        beta = -(-L // self.alpha)
        t = np.empty((beta, L + self.alpha, self.N), dtype=np.uint64)
        lib().orc_modup(self.h, L, _p(d), _p(t))
        return t

    def hadd(self, L, a, b):
        out = np.empty((2, L, self.N), dtype=np.uint64)
        lib().orc_hadd(self.h, L, _p(a), _p(b), _p(out))
        return out

    def pmult(self, L, ct, pt):
        out = np.empty((2, L, self.N), dtype=np.uint64)
        lib().orc_pmult(self.h, L, _p(ct), _p(pt), _p(out))
        return out

    def padd(self, L, ct, pt):
        out = np.empty((2, L, self.N), dtype=np.uint64)
        lib().orc_padd(self.h, L, _p(ct), _p(pt), _p(out))
        return out


# ---- seeded synthetic data: splitmix64 counter PRNG, uniform residues mod each limb
SEED = 0x486F6D756C61746F


def splitmix64(seed, n):
    with np.errstate(over="ignore"):
        i = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + i * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def uniform_limbs(moduli, N, tensor_id, lead=()):
    """uint64[*lead, len(moduli), N], limb i uniform in [0, moduli[i]) (tiny modulo bias is irrelevant here)."""
    n_lead = int(np.prod(lead)) if lead else 1
    out = np.empty((n_lead, len(moduli), N), dtype=np.uint64)
    for k in range(n_lead):
        for i, m in enumerate(moduli):
            seed = (SEED + 0x1000003 * tensor_id + 0x10001 * k + i) & 0xFFFFFFFFFFFFFFFF
            out[k, i] = splitmix64(seed, N) % np.uint64(m)
    return out.reshape(*lead, len(moduli), N)
