"""CPU model of the tcgen05 data path of block28u (csrc/block28u.cuh), checked against the library's own layout constants.

No GPU: the model re-derives the tiling (tiles, k-steps, zero chunks, Toeplitz core-matrix table, warp -> TMEM quadrant / range
mapping) from L = G * BL alone, asserts that pb200_umma_layout reports the same constants, then EXECUTES the phases the way the
hardware would be driven — operand bytes fetched through the UMMA descriptor arithmetic (start, 128 B between 8-row groups, chunk
stride between K chunks; table entry u0 + g + 2j for the B operand), accumulators per (MMA row, column), the per-warp fold of its
range — and compares every quotient-estimate digit of phase B and every low-part sum of phase C with the direct convolution
C[lane][p] = sum_k A7[lane][k] K7[p - k] folded as block28t folds it (block28.cuh, phase_mma / qhat_to_bytes / low_to_value)."""
import ctypes as C

import numpy as np
import pytest

from paillier_halo2_b200 import _lib

W = 28
SHAPES = [(8, 19, 1, 0), (8, 19, 2, 0), (8, 19, 1, 1), (16, 14, 1, 0), (16, 14, 1, 1), (16, 19, 1, 0), (16, 19, 1, 1)]


class Layout:
    """mirror of UL<C, LG, WIT> (block28u.cuh), derived independently"""

    def __init__(self, G, BL, LG):
        self.G, self.BL, self.LG = G, BL, LG
        L = self.L = G * BL
        self.K7 = 4 * L
        self.KCH = self.K7 // 16
        self.ROWS = 32 * LG
        self.CHB = 16 * self.ROWS
        self.NSH = 128 // self.ROWS
        self.SHIFTC = 16 * (self.NSH - 1)
        self.FRONT = 2 * ((self.SHIFTC + 31) // 32)
        self.BACK = self.NSH - 1
        self.KOFF = 16 * self.FRONT
        self.A_BYTES = (self.FRONT + self.KCH + self.BACK) * self.CHB
        self.RD = 3 if G == 16 else 5
        self.RCOLS = 4 * self.RD
        self.TCOLS = G * self.RCOLS
        self.DT = self.TCOLS // 4
        self.TN = (self.TCOLS - self.SHIFTC + 4 + 15) // 16 * 16
        self.RPS = G // self.NSH
        self.P_BASE_H = 4 * (L - 2)
        self.NT = {True: (4 * (L + 2) + self.TCOLS - 1) // self.TCOLS, False: (4 * L + self.TCOLS - 1) // self.TCOLS}
        self.Z0 = {h: self._z0(h) for h in (True, False)}
        self.NCM = {h: self._ncm(h) for h in (True, False)}

    def p_top(self, high, t):
        return (self.P_BASE_H + self.NT[True] * self.TCOLS if high else self.NT[False] * self.TCOLS) - 1 - self.TCOLS * t

    def p_hi(self, high, t):
        return self.p_top(high, t) - self.SHIFTC

    def k_start(self, ph):
        k_lo = max(ph - (self.TN - 1) - (self.K7 - 1), -self.SHIFTC)
        return ((k_lo + self.KOFF) // 32) * 32 - self.KOFF

    def n_ksteps(self, ph):
        return (min(ph, self.K7 - 1) - self.k_start(ph)) // 32 + 1

    def _z0(self, high):
        m = max(self.p_hi(high, t) - self.k_start(self.p_hi(high, t)) for t in range(self.NT[high]))
        return m + ((7 - m % 8) + 8) % 8

    def _ncm(self, high):
        out = 0
        for t in range(self.NT[high]):
            ph = self.p_hi(high, t)
            k_last = self.k_start(ph) + 32 * (self.n_ksteps(ph) - 1)
            out = max(out, (self.Z0[high] - ph + k_last) // 8 + self.TN // 8 + 2)
        return out


def lib_layout(G, BL, LG, wit):
    lib = _lib.load()
    out = (C.c_int32 * 20)()
    assert lib.pb200_umma_layout(G, BL, LG, wit, out) == 0
    keys = ["supported", "RD", "TCOLS", "TN", "NT_H", "NT_L", "FRONT", "BACK", "KOFF", "CHB", "A_BYTES", "Z0_H", "Z0_L", "NCM_H", "NCM_L",
            "SMEM_BYTES", "CTAS_PER_SM", "TMEM_COLS", "GAP", "P_BASE_H"]
    return dict(zip(keys, list(out)))


@pytest.mark.parametrize("G,BL,LG,wit", SHAPES)
def test_layout_constants_match_the_library(G, BL, LG, wit):
    u, m = lib_layout(G, BL, LG, wit), Layout(G, BL, LG)
    assert u["supported"] == 1
    mine = {"RD": m.RD, "TCOLS": m.TCOLS, "TN": m.TN, "NT_H": m.NT[True], "NT_L": m.NT[False], "FRONT": m.FRONT, "BACK": m.BACK, "KOFF": m.KOFF,
            "CHB": m.CHB, "A_BYTES": m.A_BYTES, "Z0_H": m.Z0[True], "Z0_L": m.Z0[False], "NCM_H": m.NCM[True], "NCM_L": m.NCM[False],
            "P_BASE_H": m.P_BASE_H}
    assert {k: u[k] for k in mine} == mine
    # resources: two TMEM buffers of TN columns; shared memory within the 227 KB a CTA may have, and CTAS_PER_SM CTAs within the SM's 228 KB
    assert 2 * m.TN <= u["TMEM_COLS"] <= 512 and u["TMEM_COLS"] * u["CTAS_PER_SM"] <= 512
    assert u["SMEM_BYTES"] <= 232448 and u["CTAS_PER_SM"] * (u["SMEM_BYTES"] + 2048) <= 233472
    if wit:
        assert u["GAP"] >= G * 32 * 8 and m.A_BYTES <= G * ((BL + 3) // 4) * 32 * 16 + u["GAP"]


def test_unsupported_shapes():
    assert lib_layout(4, 19, 1, 0)["supported"] == 0          # |n| = 1024: 9.5 k-steps per row, one range per quadrant
    assert lib_layout(16, 14, 2, 0)["supported"] == 0         # two lane groups of 16 warps do not fit an SM
    out = (C.c_int32 * 20)()
    assert _lib.load().pb200_umma_layout(5, 7, 1, 0, out) == _lib.PB200_ERR_UNSUPPORTED
    assert _lib.load().pb200_umma_layout(8, 19, 1, 0, None) == _lib.PB200_ERR_INVALID_ARG


def split7(d):
    """one signed 28-bit digit (or a slightly wider unrippled one) -> four signed 7-bit digits, the top one absorbs the remainder"""
    out = []
    for _ in range(3):
        e = ((d + 64) & 127) - 64
        out.append(e)
        d = (d - e) >> 7
    assert -128 <= d <= 127
    return out + [d]


def fold(c0, c1, c2, c3):
    v = int(c0) + (int(c1) << 7) + (int(c2) << 14) + (int(c3) << 21)
    lo = ((v + (1 << (W - 1))) & ((1 << W) - 1)) - (1 << (W - 1))
    return lo, (v - lo) >> W


def sgxt28(x):
    return ((x + (1 << 27)) & ((1 << 28) - 1)) - (1 << 27)


def run_phase(m, high, rows7, k7, rng_lanes):
    """rows7: (32 * LG, K7) s8 operand rows; k7: (K7,) constant.  Returns {(row, digit index within the phase): value} as the warps
    of block28u fold it: HIGH -> unrippled LO + CA per digit (the q-hat digit is digit - 2), LOW -> LO + CA sums (the F array)."""
    K7, CHB, TN = m.K7, m.CHB, m.TN
    # operand image in shared memory: [chunk][row][16] with zero chunks in front and behind
    img = np.zeros(m.A_BYTES + 4 * CHB, dtype=np.int64)          # + slack: reads beyond the image would be a model failure anyway
    for r in range(rows7.shape[0]):
        for k in range(K7):
            img[(m.FRONT + k // 16) * CHB + r * 16 + k % 16] = rows7[r, k]
    z0, ncm = m.Z0[high], m.NCM[high]
    cm = np.zeros(ncm * 128, dtype=np.int64)
    for u in range(ncm):
        for r in range(8):
            for b in range(16):
                d = z0 - (8 * u + r + b)
                cm[(u * 8 + r) * 16 + b] = k7[d] if 0 <= d < K7 else 0
    NT = m.NT[high]
    out = {}
    kk = np.arange(32)
    for s in range(NT):
        t = NT - 1 - s if high else s
        ph, ks0, nks = m.p_hi(high, t), m.k_start(m.p_hi(high, t)), m.n_ksteps(m.p_hi(high, t))
        assert (ks0 + m.KOFF) % 16 == 0 and (z0 - ph + ks0) % 8 == 0 and (z0 - ph + ks0) >= 0
        D = np.zeros((128, TN), dtype=np.int64)
        for ks in range(nks):
            k0 = ks0 + 32 * ks
            a_start = ((k0 + m.KOFF) // 16) * CHB
            u0 = (z0 - ph + k0) // 8
            assert u0 + TN // 8 - 1 + 2 < ncm and a_start + 15 * 128 + CHB + 8 * 16 <= m.A_BYTES
            R = np.arange(128)[:, None]
            A = img[a_start + (R // 8) * 128 + (kk[None, :] // 16) * CHB + (R % 8) * 16 + kk[None, :] % 16]        # (128, 32)
            n = np.arange(TN)[:, None]
            B = cm[(u0 + n // 8 + 2 * (kk[None, :] // 16)) * 128 + (n % 8) * 16 + kk[None, :] % 16]                 # (TN, 32)
            D += A @ B.T
        # the folds: warp cw reads TMEM quadrant q = cw % 4, i.e. MMA rows 32 q .. 32 q + 31
        for cw in range(m.G * m.LG):
            q = cw & 3
            ge, j = (0, q) if m.LG == 1 else (q & 1, q >> 1)
            ri = (m.NSH - 1 - j) * m.RPS + (cw >> 2)
            col0 = m.RCOLS * ri - m.SHIFTC + 16 * j
            assert 0 <= col0 and col0 + 4 * (m.RD + 1) <= TN
            d_top = (m.DT * (s + 1) - 1 - m.RD * ri) if high else (m.DT * (NT - s) - 1 - m.RD * ri)
            for lane in rng_lanes:
                v = D[32 * q + lane, col0:col0 + 4 * (m.RD + 1)]
                lo, ca = zip(*[fold(v[4 * e + 3], v[4 * e + 2], v[4 * e + 1], v[4 * e]) for e in range(m.RD + 1)])
                carry_g = 0
                if high and d_top == m.RD - 1:
                    t1 = lo[m.RD - 2] + ca[m.RD - 1]
                    carry_g = (t1 - sgxt28(t1)) >> W
                for e in range(m.RD):
                    key = (ge * 32 + lane, d_top - e)
                    assert key not in out
                    out[key] = lo[e] + ca[e + 1] + (carry_g if (high and e == m.RD - 3) else 0)
    return out


@pytest.mark.parametrize("G,BL,LG", [(8, 19, 1), (8, 19, 2), (16, 14, 1), (16, 19, 1)])
def test_tcgen05_data_path_model_equals_the_direct_convolution(G, BL, LG):
    m = Layout(G, BL, LG)
    L, K7 = m.L, m.K7
    rng = np.random.default_rng(100 * G + BL + LG)
    nrows = 32 * LG
    lanes = [0, 5, 31]
    # strict 28-bit digits -> s8 rows; worst-case digits on one lane
    digits = rng.integers(-(1 << 27), 1 << 27, size=(nrows, L))
    digits[5 % nrows, :] = (1 << 27) - 1
    digits[31 % nrows, ::2] = -(1 << 27)
    rows7 = np.array([[x for d in row for x in split7(int(d))] for row in digits], dtype=np.int64)
    kdig = rng.integers(-(1 << 27), 1 << 27, size=L)
    k7 = np.array([x for d in kdig for x in split7(int(d))], dtype=np.int64)
    for high in (True, False):
        got = run_phase(m, high, rows7, k7, lanes)
        p_base = m.P_BASE_H if high else 0
        nout = L + 2 if high else L
        for ge in range(LG):
            for lane in lanes:
                row = ge * 32 + lane
                a = rows7[row]
                conv = np.convolve(a, k7)                          # conv[p] = sum_k a[k] k7[p - k], p in [0, 2 K7 - 1)
                col = lambda p: int(conv[p]) if 0 <= p < len(conv) else 0  # noqa: E731
                LO, CA = {}, {-1: 0}
                for jj in range(-1, nout + 1):
                    p = p_base + 4 * jj
                    LO[jj], CA[jj + 1] = fold(col(p), col(p + 1), col(p + 2), col(p + 3))
                for jj in range(nout):
                    want = LO[jj] + CA[jj]
                    if high:
                        # block28.cuh, qhat_to_bytes: guard digit 0 takes no carry in, q-hat digit 0 (jj = 2) the overflow of guard digit 1
                        if jj == 0:
                            want = LO[0] + CA[0]                   # the model folds the columns below the phase too: same as the engine
                        if jj == 2:
                            t1 = LO[1] + CA[1]
                            want += (t1 - sgxt28(t1)) >> W
                    assert got[(row, jj)] == want, (high, row, jj)
        # every digit of the phase is produced exactly once per ciphertext, and nothing below digit 0
        for row in [ge * 32 + lane for ge in range(LG) for lane in lanes]:
            js = sorted(j for (r, j) in got if r == row)
            assert js[0] == 0 and js == list(range(js[0], js[-1] + 1)) and js[-1] >= nout - 1
