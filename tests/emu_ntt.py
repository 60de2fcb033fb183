"""Thread-level emulation (exact integers, pure Python) of the index logic of homulator_b200/csrc/ntt.cu: the 16-point
network, the column / row thread maps, the XOR-swizzled shared-memory tile, the cp.async chunk maps and the permuted
row-twiddle blob (ntt_permute_row_twiddles).  It lets the CPU test-suite prove that the kernel's data movement computes
the same transform as the oracle before any GPU time is spent; arithmetic is plain `% q` here (the FP64 modular
arithmetic itself is covered by the GPU parity tests)."""

ROW_LOG = 8
TILE = 4096


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def natural_table(psi, q, logN, inverse=False):
    """entry i = psi^(+-bitrev(i, logN)): the butterflies of stage s, group g use entry (1 << s) + g"""
    N = 1 << logN
    base = pow(psi, -1, q) if inverse else psi
    return [pow(base, bitrev(i, logN), q) for i in range(N)]


def permute_row_twiddles(nat, logN):
    N = 1 << logN
    R1 = N >> ROW_LOG
    out = [0] * N
    for r in range(R1):
        o = (r // 16) * TILE
        rr, base = r % 16, R1 + r
        for s in range(4):
            for g in range(1 << s):
                out[o + rr * 16 + (1 << s) + g] = nat[(base << s) + g]
        for l in range(16):
            ts = rr * 16 + l
            out[o + 256 + ts] = nat[(base << 4) + l]
            for x in range(2):
                out[o + 512 + ts * 2 + x] = nat[(base << 5) + 2 * l + x]
            for k in range(2):
                for x in range(2):
                    out[o + 1024 + k * 512 + ts * 2 + x] = nat[(base << 6) + 4 * l + 2 * k + x]
            for k in range(4):
                for x in range(2):
                    out[o + 2048 + k * 512 + ts * 2 + x] = nat[(base << 7) + 8 * l + 2 * k + x]
    return out


def ct_level(a, T, w, q):
    D = 8 >> T
    for g in range(1 << T):
        for o in range(D):
            i, j = g * 2 * D + o, g * 2 * D + o + D
            t = a[j] * w[g] % q
            a[i], a[j] = (a[i] + t) % q, (a[i] - t) % q


def gs_level(a, T, w, q):
    D = 8 >> T
    for g in range(1 << T):
        for o in range(D):
            i, j = g * 2 * D + o, g * 2 * D + o + D
            a[i], a[j] = (a[i] + a[j]) % q, (a[i] - a[j]) * w[g] % q


# ------------------------------------------------------------------------------------------------ column passes
def col_tile(data, logN, tw, q, tile, inverse, post=1, NT=256):
    """one work item of ntt_fwd_cols / ntt_inv_cols on the limb `data` (list of N ints), in place"""
    LOGR1 = logN - ROW_LOG
    R1 = 1 << LOGR1
    CT = 16 * NT  # points per column-pass tile
    C = CT // R1
    SKIP = 8 - LOGR1
    col0 = tile * C
    sm = [None] * CT
    # col_issue: chunk qd = tid + 256k -> row qd / (C/2), column pair qd % (C/2); linear shared-memory layout
    for tid in range(NT):
        for k in range(8):
            qd = tid + NT * k
            row, cc = qd // (C // 2), qd % (C // 2)
            for x in range(2):
                sm[2 * qd + x] = data[row * 256 + col0 + 2 * cc + x]
    assert all(v is not None for v in sm)
    regs = {}
    if not inverse:
        for tid in range(NT):
            a = [sm[tid + NT * j] for j in range(16)]
            for T in range(4):
                ct_level(a, T, tw[(1 << T):(2 << T)], q)
            regs[tid] = a
        for tid in range(NT):
            for j in range(16):
                sm[tid + NT * j] = regs[tid][j]
        for tid in range(NT):
            c, u = tid % C, tid // C
            a = [sm[(16 * u + j) * C + c] for j in range(16)]
            for T in range(max(SKIP, 0), 4):
                b = (1 << (LOGR1 - 4 + T)) + (u << T)
                ct_level(a, T, tw[b:b + (1 << T)], q)
            for j in range(16):
                data[(16 * u + j) * 256 + col0 + c] = a[j]
    else:
        for tid in range(NT):
            c, u = tid % C, tid // C
            a = [sm[(16 * u + j) * C + c] for j in range(16)]
            for T in range(3, max(SKIP, 0) - 1, -1):
                b = (1 << (LOGR1 - 4 + T)) + (u << T)
                gs_level(a, T, tw[b:b + (1 << T)], q)
            regs[tid] = a
        for tid in range(NT):
            c, u = tid % C, tid // C
            for j in range(16):
                sm[(16 * u + j) * C + c] = regs[tid][j]
        for tid in range(NT):
            c, u = tid % C, tid // C
            a = [sm[tid + NT * j] for j in range(16)]
            for T in range(3, -1, -1):
                gs_level(a, T, tw[(1 << T):(2 << T)], q)
            for j in range(16):
                data[(u + (R1 // 16) * j) * 256 + col0 + c] = a[j] * post % q


# ------------------------------------------------------------------------------------------------ row passes
def row_addr(lane, warp):
    l, rr = lane & 15, 2 * warp + (lane >> 4)
    base = rr * 2048
    xa, xb, xc = [], [], []
    for m in range(8):
        xa.append(base + ((((l >> 1) ^ m) << 4) | ((l & 1) << 3)))
        xb.append(base + l * 128 + ((m ^ (l & 7)) << 4))
        g = 2 * m + (l >> 3)
        xc.append(base + g * 128 + (((l & 7) ^ (g & 7)) << 4))
    return xa, xb, xc


def bank_conflict_degree(byte_addrs, width):
    """extra shared-memory wavefronts of one warp request (0 = conflict-free).  8-byte accesses are served per
    half-warp, 16-byte accesses per quarter-warp; inside a phase two lanes conflict when they touch the same 4-byte
    bank at different addresses."""
    per_phase = 128 // width
    extra = 0
    for p0 in range(0, len(byte_addrs), per_phase):
        banks = {}
        for a in byte_addrs[p0:p0 + per_phase]:
            for w in range(a // 4, (a + width) // 4):
                banks.setdefault(w % 32, set()).add(w)
        extra += max(len(v) for v in banks.values()) - 1
    return extra


def row_tile(data, logN, blob_all, q, tile, inverse, check_banks=False):
    """one item of ntt_rows<INV> on CTA tile `tile` (16 rows) of the limb `data`, in place"""
    blob = blob_all[tile * TILE:(tile + 1) * TILE]
    base = tile * TILE
    sm = {}
    # row_issue: warp-local, chunk qd = lane + 32k of the warp's 4 KB, swizzled destination
    for warp in range(8):
        for lane in range(32):
            for k in range(8):
                qd = lane + 32 * k
                g, cpos = (qd >> 3) & 15, qd & 7
                dst = (2 * warp + (qd >> 7)) * 2048 + g * 128 + ((cpos ^ (g & 7)) << 4)
                for x in range(2):
                    assert dst + 8 * x not in sm
                    sm[dst + 8 * x] = data[base + warp * 512 + 2 * qd + x]
    assert len(sm) == TILE
    for warp in range(8):
        addrs = [row_addr(lane, warp) for lane in range(32)]
        if check_banks:
            for j in range(16):
                assert bank_conflict_degree([addrs[ln][0][j & 7] + 128 * j for ln in range(32)], 8) == 0, ("xa", j)
            for m in range(8):
                assert bank_conflict_degree([addrs[ln][1][m] for ln in range(32)], 16) == 0, ("xb", m)
                assert bank_conflict_degree([addrs[ln][2][m] for ln in range(32)], 16) == 0, ("xc", m)
        regs = []
        if not inverse:
            for lane in range(32):
                xa, xb, xc = addrs[lane]
                tid, rr = warp * 32 + lane, 2 * warp + (lane >> 4)
                a = [sm[xa[j & 7] + 128 * j] for j in range(16)]
                for T in range(4):
                    ct_level(a, T, blob[rr * 16 + (1 << T): rr * 16 + (2 << T)], q)
                regs.append(a)
            for lane in range(32):
                xa = addrs[lane][0]
                for j in range(16):
                    sm[xa[j & 7] + 128 * j] = regs[lane][j]
            regs = []
            for lane in range(32):
                xb = addrs[lane][1]
                tid = warp * 32 + lane
                a = []
                for m in range(8):
                    a += [sm[xb[m]], sm[xb[m] + 8]]
                ct_level(a, 0, [blob[256 + tid]], q)
                ct_level(a, 1, blob[512 + 2 * tid: 512 + 2 * tid + 2], q)
                ct_level(a, 2, blob[1024 + 2 * tid: 1024 + 2 * tid + 2] + blob[1536 + 2 * tid: 1536 + 2 * tid + 2], q)
                w = []
                for m in range(4):
                    w += blob[2048 + 512 * m + 2 * tid: 2048 + 512 * m + 2 * tid + 2]
                ct_level(a, 3, w, q)
                regs.append(a)
            for lane in range(32):
                xb = addrs[lane][1]
                for m in range(8):
                    sm[xb[m]], sm[xb[m] + 8] = regs[lane][2 * m], regs[lane][2 * m + 1]
            for lane in range(32):
                xc = addrs[lane][2]
                l16, rr = lane & 15, 2 * warp + (lane >> 4)
                o = base + rr * 256 + (l16 >> 3) * 16 + (l16 & 7) * 2
                for m in range(8):
                    data[o + 32 * m], data[o + 32 * m + 1] = sm[xc[m]], sm[xc[m] + 8]
        else:
            for lane in range(32):
                xb = addrs[lane][1]
                tid = warp * 32 + lane
                a = []
                for m in range(8):
                    a += [sm[xb[m]], sm[xb[m] + 8]]
                w = []
                for m in range(4):
                    w += blob[2048 + 512 * m + 2 * tid: 2048 + 512 * m + 2 * tid + 2]
                gs_level(a, 3, w, q)
                gs_level(a, 2, blob[1024 + 2 * tid: 1024 + 2 * tid + 2] + blob[1536 + 2 * tid: 1536 + 2 * tid + 2], q)
                gs_level(a, 1, blob[512 + 2 * tid: 512 + 2 * tid + 2], q)
                gs_level(a, 0, [blob[256 + tid]], q)
                regs.append(a)
            for lane in range(32):
                xb = addrs[lane][1]
                for m in range(8):
                    sm[xb[m]], sm[xb[m] + 8] = regs[lane][2 * m], regs[lane][2 * m + 1]
            for lane in range(32):
                xa = addrs[lane][0]
                l16, rr = lane & 15, 2 * warp + (lane >> 4)
                a = [sm[xa[j & 7] + 128 * j] for j in range(16)]
                for T in range(3, -1, -1):
                    gs_level(a, T, blob[rr * 16 + (1 << T): rr * 16 + (2 << T)], q)
                for j in range(16):
                    data[base + rr * 256 + l16 + 16 * j] = a[j]


def forward(x, psi, q, logN, NT=256):
    nat = natural_table(psi, q, logN)
    rows = permute_row_twiddles(nat, logN)
    d = list(x)
    C = (16 * NT) >> (logN - ROW_LOG)
    for tile in range(256 // C):
        col_tile(d, logN, nat, q, tile, False, NT=NT)
    for tile in range((1 << logN) // TILE):
        row_tile(d, logN, rows, q, tile, False, check_banks=(tile == 0))
    return d


def inverse(x, psi, q, logN, NT=256):
    nat = natural_table(psi, q, logN, inverse=True)
    rows = permute_row_twiddles(nat, logN)
    d = list(x)
    C = (16 * NT) >> (logN - ROW_LOG)
    for tile in range((1 << logN) // TILE):
        row_tile(d, logN, rows, q, tile, True, check_banks=(tile == 0))
    ninv = pow(1 << logN, -1, q)
    for tile in range(256 // C):
        col_tile(d, logN, nat, q, tile, True, post=ninv, NT=NT)
    return d
