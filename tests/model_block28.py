"""Limb-exact Python model of the block28 CUDA engine (paillier_halo2_b200/csrc/block28.cuh).

Not the oracle: this file models OUR kernel's arithmetic (radix-2^28 signed digits, G x BL block
products, lazy Barrett with truncated high product) so its bounds can be fuzzed on the CPU.  Every
64-bit accumulator is range-checked against the signed 64-bit limits the kernel relies on.
"""
from __future__ import annotations

W = 28
M = 1 << W
H = 1 << (W - 1)
I64 = 1 << 63


def sgxt(v: int) -> int:
    """centered residue mod 2^28 in [-2^27, 2^27)"""
    return ((v + H) & (M - 1)) - H


class Params:
    def __init__(self, G: int, BL: int, margin: int = 10):
        self.G, self.BL = G, BL
        self.L = G * BL
        self.beta = 14 * (2 * self.L - 1)          # 2*beta = s1 + s2 = 28(2L-1)
        self.kN = self.beta - margin               # bit length of the normalised modulus Nt
        self.margin = margin

    def key(self, N: int):
        assert N % 2 == 1 and N.bit_length() <= self.kN
        sh = self.kN - N.bit_length()
        Nt = N << sh
        mu = (1 << (2 * self.beta)) // Nt
        assert mu.bit_length() <= W * self.L - 1
        return sh, Nt, mu


def to_digits(v: int, L: int):
    """exact integer -> L strict centered digits (value must fit)"""
    out = []
    for _ in range(L):
        d = sgxt(v)
        out.append(d)
        v = (v - d) >> W
    assert v == 0, "value does not fit"
    return out


def from_digits(d):
    return sum(x << (W * i) for i, x in enumerate(d))


def block(d, b, BL):
    return d[b * BL:(b + 1) * BL]


def antidiag(P, A, B, pairs, mode_of=None):
    """D = sum over block pairs (i,j) of A_i * B_j (column sums, 2BL-1 signed 64-bit accumulators)."""
    BL = P.BL
    acc = [0] * (2 * BL - 1)
    for (i, j, mode, dbl) in pairs:
        a, b = block(A, i, BL), block(B, j, BL)
        for x in range(BL):
            for y in range(BL):
                if mode == "F" or (mode == "LT" and x + y <= BL - 1) or (mode == "UTG" and x + y >= BL - 2):
                    acc[x + y] += (2 if dbl else 1) * a[x] * b[y]
                elif mode == "SQ":
                    if x == y:
                        acc[x + y] += a[x] * b[y]
                    elif x < y:
                        acc[x + y] += 2 * a[x] * b[y]
    for c in acc:
        assert -I64 < c < I64, "64-bit accumulator overflow"
    return acc


def normalize_D(P, acc):
    """ripple the 2BL-1 column sums into 2BL strict digits + one small spill digit (weight 2^(28*2BL))"""
    out = []
    carry = 0
    for c in acc:
        t = c + carry
        assert -I64 < t < I64
        d = sgxt(t)
        out.append(d)
        carry = (t - d) >> W
    d = sgxt(carry)
    out.append(d)                 # digit 2BL-1
    spill = (carry - d) >> W
    assert abs(spill) <= 4, "spill digit larger than expected"
    out.append(spill)             # digit 2BL
    return out


def sched_full(G):
    """role t -> [(d, [(i, j)])]: cyclic antidiagonals t and t+G of a full product"""
    out = []
    for t in range(G):
        lo = [(i, t - i) for i in range(0, t + 1)]
        hi = [(i, t + G - i) for i in range(t + 1, G)]
        out.append([(t, lo), (t + G, hi)])
    return out


def product(P, A, B, nblocks_out, which, sqr=False):
    """X blocks [0, nblocks_out) of A*B.  which: "full" | "high" (blocks >= G exact up to the guard) | "low" (blocks < G)"""
    G, BL = P.G, P.BL
    X = [0] * (nblocks_out * BL)
    Ds = {}
    for d in range(2 * G - 1):
        pairs = []
        for i in range(G):
            j = d - i
            if not (0 <= j < G):
                continue
            if which == "full":
                if sqr:
                    if i < j:
                        pairs.append((i, j, "F", True))
                    elif i == j:
                        pairs.append((i, j, "SQ", False))
                else:
                    pairs.append((i, j, "F", False))
            elif which == "high":
                if d >= G:
                    pairs.append((i, j, "F", False))
                elif d == G - 1:
                    pairs.append((i, j, "UTG", False))
            elif which == "low":
                if d < G - 1:
                    pairs.append((i, j, "F", False))
                elif d == G - 1:
                    pairs.append((i, j, "LT", False))
        if not pairs:
            continue
        D = normalize_D(P, antidiag(P, A, B, pairs))
        Ds[d] = D
    # step 1: Lo parts (plain stores); step 2: Hi parts added with a per-block ripple; step 3: deposits
    nb = nblocks_out
    for d, D in Ds.items():
        if d < nb:
            for k in range(BL):
                X[d * BL + k] += D[k]
    carries = {}
    for d, D in Ds.items():
        if d + 1 < nb:
            carry = 0
            for k in range(BL):
                t = X[(d + 1) * BL + k] + D[BL + k] + carry
                assert abs(t) < (1 << 31)
                dd = sgxt(t)
                X[(d + 1) * BL + k] = dd
                carry = (t - dd) >> W
            carries[d] = carry
    for d, D in Ds.items():
        if d + 2 < nb:
            X[(d + 2) * BL] += carries.get(d, 0) + D[2 * BL]
    return X


def ripple_blocks(P, X, nblocks):
    """per-block ripple: strict digits, block-top keeps the carry (loose)"""
    BL = P.BL
    out = list(X)
    for b in range(nblocks):
        carry = 0
        for k in range(BL):
            t = out[b * BL + k] + carry
            if k < BL - 1:
                d = sgxt(t)
                out[b * BL + k] = d
                carry = (t - d) >> W
            else:
                out[b * BL + k] = t
                assert abs(t) < (1 << 31)
    return out


def mulmod(P, keyc, A, B, sqr=False):
    """lazy Barrett: returns R (L digits, strict with loose block tops), R == A*B (mod Nt), |R| < 2^beta"""
    sh, Nt, mu = keyc
    G, BL, L = P.G, P.BL, P.L
    X = product(P, A, B, 2 * G, "full", sqr)                     # 2L digits, loose (sum of two D digits)
    # q1 = X digits [L-1, 2L-1), with digit 2L-1 folded into 2L-2
    q1 = X[L - 1:2 * L - 1]
    q1[L - 1] += X[2 * L - 1] << W
    assert abs(q1[L - 1]) < (1 << 31)
    mu_d = keyc_digits(P, keyc)[1]
    Y = product(P, q1, mu_d, 2 * G, "high")
    qh = Y[L:2 * L]                                              # q-hat digits (loose)
    Nt_d = keyc_digits(P, keyc)[0]
    Pl = product(P, qh, Nt_d, G, "low")
    R = [X[i] - Pl[i] for i in range(L)]
    out = [0] * L
    for b in range(G):                                           # per-block ripple, carry deposited into the next block
        carry = 0
        for k in range(BL):
            t = R[b * BL + k] + carry
            assert abs(t) < (1 << 31)
            dd = sgxt(t)
            out[b * BL + k] = dd
            carry = (t - dd) >> W
        if b + 1 < G:
            R[(b + 1) * BL] += carry                             # (the kernel does this after a barrier)
    return out                                                   # carry out of the top block dropped: mod 2^(28L)


_cache = {}


def keyc_digits(P, keyc):
    k = (P.L, keyc[1])
    if k not in _cache:
        _cache[k] = (to_digits(keyc[1], P.L), to_digits(keyc[2], P.L))
    return _cache[k]


def canonical(P, keyc, R, N):
    """final step: (R * 2^sh mod Nt) >> sh  ==  R mod N"""
    sh, Nt, mu = keyc
    two_sh = to_digits(1 << sh, P.L)
    Y = mulmod(P, keyc, R, two_sh)
    v = from_digits(Y)
    assert abs(v) < (1 << P.beta)
    v %= Nt
    assert v % (1 << sh) == 0
    return v >> sh


# ---------------------------------------------------------------------------------------------
# IMMA variant of phases B and C: each signed 28-bit digit splits carry-free into four signed 7-bit digits
# (s8 operands of mma.sync m16n8k32), the constant is a Toeplitz operand, s32 columns are folded back.

def split7(d):
    """28-bit digit (|d| <= 2^27 + small) -> 4 signed 7-bit digits, the top one absorbs the remainder"""
    out = []
    for _ in range(3):
        e = ((d + 64) & 127) - 64
        out.append(e)
        d = (d - e) >> 7
    out.append(d)
    assert all(-128 <= e <= 127 for e in out), out
    return out


def digits7(D):
    out = []
    for d in D:
        out.extend(split7(d))
    return out


def conv_columns(A7, B7, p_lo, p_hi):
    """s32 column sums c_p = sum_k A7[k] * B7[p-k] for p in [p_lo, p_hi)"""
    n = len(B7)
    cols = []
    for p in range(p_lo, p_hi):
        s = 0
        for k in range(max(0, p - n + 1), min(len(A7), p + 1)):
            s += A7[k] * B7[p - k]
        assert -(1 << 31) < s < (1 << 31), "s32 accumulator overflow"
        cols.append(s)
    return cols


def fold28(cols):
    """groups of 4 radix-2^7 columns -> (lo, carry) per 28-bit digit; digit j = lo[j] + carry[j-1]"""
    lo, carry = [], []
    for j in range(len(cols) // 4):
        v = cols[4 * j] + (cols[4 * j + 1] << 7) + (cols[4 * j + 2] << 14) + (cols[4 * j + 3] << 21)
        l = sgxt(v)
        lo.append(l)
        carry.append((v - l) >> W)
    return lo, carry


def mulmod_imma(P, keyc, A, B, sqr=False):
    sh, Nt, mu = keyc
    G, BL, L = P.G, P.BL, P.L
    X = product(P, A, B, 2 * G, "full", sqr)
    q1 = X[L - 1:2 * L - 1]
    q1[L - 1] += X[2 * L - 1] << W
    Nt_d, mu_d = keyc_digits(P, keyc)
    # phase B: digits j >= L of q1*mu, two guard digits below
    cols = conv_columns(digits7(q1), digits7(mu_d), 4 * (L - 2), 4 * 2 * L)
    lo, carry = fold28(cols)                      # index jj = j - (L-2)
    # q-hat digits are NOT rippled: lo + carry-in of the fold is already within the range the s8 split absorbs
    # (|d| <= 2^27 + 2^17, top 7-bit piece within [-65, 64]); only the two guard digits feed a carry into digit 0
    qh = []
    c = 0
    for jj in range(2):
        t = lo[jj] + (carry[jj - 1] if jj else 0) + c
        c = (t - sgxt(t)) >> W
    for jj in range(2, len(lo)):
        d = lo[jj] + carry[jj - 1] + (c if jj == 2 else 0)
        assert abs(d) < (1 << W) - (1 << 21)
        qh.append(d)
    # phase C: low L digits of qh*Nt (exact)
    cols = conv_columns(digits7(qh), digits7(Nt_d), 0, 4 * L)
    lo, carry = fold28(cols)
    out = []
    c = 0
    for j in range(L):
        t = X[j] - lo[j] - (carry[j - 1] if j else 0) + c
        d = sgxt(t)
        c = (t - d) >> W
        out.append(d)
    return out


# ---------------------------------------------------------------------------------------------
# Witness step (block28.cuh: w_tail / mulmod_w): exact (q, rem) of one mul_mod on canonical operands.
# Chain values are kept as strict digits of x * 2^s, 2s = sh_w, Nt_w = n^2 << sh_w.

def witness_key(P, N: int):
    """per-key constants of the witness engine: sh_w even"""
    sh = P.kN - N.bit_length()
    if sh & 1:
        sh -= 1
    Nt = N << sh
    mu = (1 << (2 * P.beta)) // Nt
    assert mu.bit_length() <= W * P.L - 1
    inv = float(1 << 20) / float(Nt >> (W * (P.L - 2) - 20))
    return (sh, Nt, mu), inv


def _imma_parts(P, keyc, A, B, sqr):
    """phases A, B, C of mulmod_imma, returning T digits, q-hat digits and the raw (unrippled) digits of V'"""
    L = P.L
    X = product(P, A, B, 2 * P.G, "full", sqr)
    q1 = X[L - 1:2 * L - 1]
    q1[L - 1] += X[2 * L - 1] << W
    Nt_d, mu_d = keyc_digits(P, keyc)
    cols = conv_columns(digits7(q1), digits7(mu_d), 4 * (L - 2), 4 * 2 * L)
    lo, carry = fold28(cols)
    qh, c = [], 0
    for jj in range(2):
        t = lo[jj] + (carry[jj - 1] if jj else 0) + c
        c = (t - sgxt(t)) >> W
    for jj in range(2, len(lo)):
        qh.append(lo[jj] + carry[jj - 1] + (c if jj == 2 else 0))      # not rippled (see mulmod_imma)
    qh = qh[:L]
    cols = conv_columns(digits7(qh), digits7(Nt_d), 0, 4 * L)
    lo, carry = fold28(cols)
    raw = [X[j] - lo[j] - (carry[j - 1] if j else 0) for j in range(L)]
    return qh, raw


def _resolve(P, digs, couts):
    """carries between blocks through flags, one round per barrier; the top block's carry leaves (mod 2^(28L))"""
    G, BL = P.G, P.BL
    rounds = 0
    while True:
        couts[G - 1] = 0
        rounds += 1
        if not any(couts):
            return rounds
        cins = [0] + couts[:G - 1]
        couts = [0] * G
        for b in range(G):
            c = cins[b]
            if c:
                for k in range(BL):
                    t = digs[b * BL + k] + c
                    digs[b * BL + k] = t & (M - 1)
                    c = t >> W
                couts[b] = c


def extract64(F, L, bit):
    p, off = divmod(bit, W)
    w = 0
    for i in range(4):
        d = F[p + i] if p + i < L else 0
        sft = W * i - off
        if i == 0:
            w = d >> off
        elif sft < 64:
            w |= (d << sft) & ((1 << 64) - 1)
    return w


def witness_step(P, keyw, inv, A, B, words_out, sqr=False, stats=None):
    """A, B: strict digits of a*2^s, b*2^s.  Returns (q, rem, next) with next = strict digits of rem*2^s."""
    import math
    sh, Nt, mu = keyw
    G, BL, L = P.G, P.BL, P.L
    qd, rd = _imma_parts(P, keyw, A, B, sqr)
    ntu = [(Nt >> (W * p)) & (M - 1) for p in range(L)]
    # k estimate from the strict top two digits of the top block
    c, top = 0, []
    for k in range(BL):
        t = rd[(G - 1) * BL + k] + c
        d = sgxt(t)
        c = (t - d) >> W
        top.append(d)
    adj = max(-3, min(3, math.floor((top[BL - 1] * float(M) + top[BL - 2]) * inv)))
    first, passes = True, 0
    while True:
        passes += 1
        if first or adj:
            cR, cQ = [0] * G, [0] * G
            for b in range(G):
                c = 0
                for k in range(BL):
                    t = rd[b * BL + k] - adj * ntu[b * BL + k] + c
                    assert abs(t) < (1 << 31)
                    rd[b * BL + k] = t & (M - 1)
                    c = t >> W
                cR[b] = c
                c = adj if b == 0 else 0
                for k in range(BL):
                    t = qd[b * BL + k] + c
                    qd[b * BL + k] = t & (M - 1)
                    c = t >> W
                cQ[b] = c
            r1 = _resolve(P, rd, cR)
            r2 = _resolve(P, qd, cQ)
            if stats is not None:
                stats["rounds"] = max(stats.get("rounds", 0), r1, r2)
        first = False
        c = 0
        for b in range(G - 1, -1, -1):
            f = 0
            for k in range(BL):
                if rd[b * BL + k] != ntu[b * BL + k]:
                    f = 1 if rd[b * BL + k] > ntu[b * BL + k] else -1
            if b == G - 1 and rd[L - 1] >= H:
                f = -2
            if c == 0:
                c = f
        adj = -1 if c == -2 else (1 if c >= 0 else 0)
        if not adj:
            break
        assert passes < 8
    if stats is not None:
        stats["passes"] = max(stats.get("passes", 0), passes)
    q = sum(extract64(qd, L, 64 * j) << (64 * j) for j in range(words_out))
    rem = sum(extract64(rd, L, 64 * j + sh) << (64 * j) for j in range(words_out))
    s = sh >> 1
    pd, off = divmod(s, W)
    nxt, c = [], 0
    for p0 in range(L):
        p = p0 + pd
        lo = rd[p] if p < L else 0
        hi = rd[p + 1] if p + 1 < L else 0
        x = ((lo >> off) | ((hi << (W - off)) if off else 0)) & (M - 1)
        t = x + c
        d = sgxt(t)
        c = (t - d) >> W
        nxt.append(d)
    assert c == 0
    return q, rem, nxt
