"""Op-sequence replay (BASELINE.json configs[4]).

The reference simulates ONE operation per run and explicitly does not chain them ("NotSuppotr the continuous operation
simulate", reference src/Operation.cpp:636,675,714); its repository contains no application trace.  This module defines
a small trace format over the five operations the reference's CLI accepts (reference bench_test/bench_micro24.cpp:29-48)
and replays it on real data through the C ABI, one kernel schedule per op, all on one stream.

A trace is a list of tuples:
    ("hrotate", dst, src, r)        dst = rotate(src, 5^r)            (rotation key index r)
    ("pmult",   dst, src, pt)       dst = src * plaintext[pt]
    ("hadd",    dst, a, b)          dst = a + b
    ("padd",    dst, src, pt)       dst = src + plaintext[pt]
    ("hmult",   dst, a, b)          dst = rescale(relin(a * b))       (drops one level)
`bsgs_trace(n1, n2)` is the synthetic rotation-heavy sequence used as configs[4]: a baby-step/giant-step
matrix-vector product (the inner loop of CKKS bootstrapping's CoeffToSlot): n1 baby rotations, n1*n2 plaintext
multiplications and additions, n2 giant rotations, then one hmult.
"""
from collections import Counter


def bsgs_trace(n1=4, n2=4):
    t = []
    for i in range(1, n1):                       # baby steps: rot_i = rotate(x, i)
        t.append(("hrotate", "b%d" % i, "x", i))
    for j in range(n2):                          # giant steps
        acc = None
        for i in range(n1):
            src = "x" if i == 0 else "b%d" % i
            if acc is None:
                acc = "g%d" % j
                t.append(("pmult", acc, src, j * n1 + i))
            else:
                t.append(("pmult", "m", src, j * n1 + i))
                t.append(("hadd", acc, acc, "m"))
        if j:
            t.append(("hrotate", acc, acc, n1 * j))
            t.append(("hadd", "y", "y", acc))
        else:
            t.append(("hadd", "y", acc, acc))
    t.append(("hmult", "z", "y", "y"))
    return t


def trace_counts(trace):
    return dict(Counter(op[0] for op in trace))


def replay(ctx, L, trace, x, plaintexts, rot_keys, evk):
    """Run `trace` at level L.  x: ct [2][L][N]; plaintexts: dict idx -> [L][N]; rot_keys: dict r -> key tensor;
    evk: relinearisation key.  Returns the dict of named ciphertexts (device tensors)."""
    N2 = 2 * ctx.N
    env = {"x": x}
    for op in trace:
        kind, dst = op[0], op[1]
        if kind == "hrotate":
            env[dst] = ctx.hrotate(L, env[op[2]], rot_keys[op[3]], pow(5, op[3], N2))
        elif kind == "pmult":
            env[dst] = ctx.pmult(L, env[op[2]], plaintexts[op[3]])
        elif kind == "padd":
            env[dst] = ctx.padd(L, env[op[2]], plaintexts[op[3]])
        elif kind == "hadd":
            env[dst] = ctx.hadd(L, env[op[2]], env[op[3]])
        elif kind == "hmult":
            env[dst] = ctx.hmult(L, env[op[2]], env[op[3]], evk)
        else:
            raise ValueError("unknown op in trace: %r" % (kind,))
    return env


def replay_sharded(sh, trace, x_own, plaintexts2, rot_keys_own, evk_own):
    """The same trace on limb-sharded operands, one rank per GPU (homulator_b200.api.ShardP2P, peer-direct exchanges over
    NVLink).  x_own: this rank's limbs of the input [2][nq][N]; plaintexts2: idx -> [2][nq][N] (owned limbs, repeated for
    both components); rot_keys_own / evk_own: this rank's key slices [beta][2][n_own_ext][N].  The trace must end with its
    only hmult (every op runs at the level `sh` was built for)."""
    N2 = 2 * sh.ctx.N
    env = {"x": x_own}
    for op in trace:
        kind, dst = op[0], op[1]
        if kind == "hrotate":
            env[dst] = sh.hrotate(env[op[2]], rot_keys_own[op[3]], pow(5, op[3], N2))
        elif kind == "pmult":
            env[dst] = sh.pmult(env[op[2]], plaintexts2[op[3]])
        elif kind == "padd":
            env[dst] = sh.padd(env[op[2]], plaintexts2[op[3]])
        elif kind == "hadd":
            env[dst] = sh.hadd(env[op[2]], env[op[3]])
        elif kind == "hmult":
            env[dst] = sh.hmult(env[op[2]], env[op[3]], evk_own)
        else:
            raise ValueError("unknown op in trace: %r" % (kind,))
    return env
