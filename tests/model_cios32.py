"""Word-exact Python model of fr_to_mont32 (paillier_halo2_b200/csrc/cells.cu): x * 2^256 mod p for BN254 Fr as a CIOS in four 64-bit
steps over two arrays of 32-bit words whose 64-bit pairs sit at even (E) and odd (O) offsets, every row of four 32 x 32 products one
carry chain followed by three carry words.  The model asserts that no carry is ever lost; tests/test_model_witness.py runs it."""
import random
P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R = 1 << 256
R2 = (R * R) % P
INV64 = (-pow(P, -1, 1 << 64)) % (1 << 64)
M32 = (1 << 32) - 1
def words32(v, n): return [(v >> (32 * i)) & M32 for i in range(n)]
p32 = words32(P, 8); r32 = words32(R2, 8)
NW = 20
class Arr:
    """array of 32-bit words; pair(k) = 64-bit value at words (k, k+1)"""
    def __init__(self): self.w = [0] * NW
def chain(arr, base, a, b_words):
    """arr pairs at offsets base, base+2, ...: += a * b_words[t]; carry chained; final carry propagated through 2 more words"""
    carry = 0
    k = base
    for b in b_words:
        v = arr.w[k] + (arr.w[k + 1] << 32) + a * b + carry
        arr.w[k] = v & M32; arr.w[k + 1] = (v >> 32) & M32; carry = v >> 64
        k += 2
    # propagate the carry: addc.cc, addc.cc, addc (the last word is a carry-count word: must not overflow)
    for t in range(3):
        v = arr.w[k + t] + carry
        arr.w[k + t] = v & M32; carry = v >> 32
    assert carry == 0, "carry lost"
def row(E, O, base, a, b):
    """T += a(64-bit) * b(8 x 32-bit words) at 32-bit offset base (even)"""
    a0, a1 = a & M32, a >> 32
    chain(E, base, a0, b[0::2])          # offsets base, +2, +4, +6
    chain(O, base + 1, a0, b[1::2])      # offsets base+1, +3, +5, +7
    chain(O, base + 1, a1, b[0::2])      # a1*b_even at base+1+j
    chain(E, base + 2, a1, b[1::2])      # a1*b_odd at base+1+j (even offsets base+2..)
def to_mont(x):
    E, O = Arr(), Arr()
    xw = [(x >> (64 * i)) & ((1 << 64) - 1) for i in range(4)]
    cin = 0
    for i in range(4):
        if xw[i]: row(E, O, 2 * i, xw[i], r32)
        s0 = E.w[2 * i] + O.w[2 * i] + cin
        s1 = E.w[2 * i + 1] + O.w[2 * i + 1] + (s0 >> 32)
        Ti = (s0 & M32) | ((s1 & M32) << 32)
        m = (Ti * INV64) & ((1 << 64) - 1)
        row(E, O, 2 * i, m, p32)
        s0 = E.w[2 * i] + O.w[2 * i] + cin
        s1 = E.w[2 * i + 1] + O.w[2 * i + 1] + (s0 >> 32)
        assert (s0 & M32) == 0 and (s1 & M32) == 0
        cin = s1 >> 32
    # result = words 8..  (E + O + cin)
    res = 0
    for k in range(8, NW):
        res += (E.w[k] + O.w[k]) << (32 * (k - 8))
    res += cin
    if res >= P: res -= P
    assert res < P
    return res


def k_words(nx):
    return words32(pow(2, 256 + 64 * nx, P), 8)


def redc_short(x, nx):
    """fr_redc<NX> of cells.cu: x < 2^(64 nx); T = x * K_nx (nx rows), nx reduction rows with radix 2^(64 nx), result T / 2^(64 nx) < 2p,
    one conditional subtraction.  Same E/O arrays, same chains, every carry asserted."""
    assert 0 <= x < 1 << (64 * nx)
    E, O = Arr(), Arr()
    kw = k_words(nx)
    xw = [(x >> (64 * i)) & ((1 << 64) - 1) for i in range(nx)]
    for i in range(nx):
        row(E, O, 2 * i, xw[i], kw)
    cin = 0
    for i in range(nx):
        s0 = E.w[2 * i] + O.w[2 * i] + cin
        s1 = E.w[2 * i + 1] + O.w[2 * i + 1] + (s0 >> 32)
        Ti = (s0 & M32) | ((s1 & M32) << 32)
        m = (Ti * INV64) & ((1 << 64) - 1)
        row(E, O, 2 * i, m, p32)
        s0 = E.w[2 * i] + O.w[2 * i] + cin
        s1 = E.w[2 * i + 1] + O.w[2 * i + 1] + (s0 >> 32)
        assert (s0 & M32) == 0 and (s1 & M32) == 0
        cin = s1 >> 32
    res = cin
    for k in range(8):          # the kernel reads exactly four 64-bit pairs of each array
        res += (E.w[2 * nx + k] + O.w[2 * nx + k]) << (32 * k)
    assert all(w == 0 for w in E.w[2 * nx + 8:]) and all(w == 0 for w in O.w[2 * nx + 8:]), "value above the four result words"
    assert res < 2 * P
    if res >= P: res -= P
    return res


def self_test(count=3000, seed=5):
    rng = random.Random(seed)
    tests = [0, 1, 2, (1 << 64) - 1, (1 << 64), (1 << 128) - 1, (1 << 135) - 1, (1 << 136) - 1, P - 1]
    tests += [rng.getrandbits(rng.choice([15, 64, 72, 128, 135, 136, 200, 253])) for _ in range(count)]
    for x in tests:
        x %= P
        assert to_mont(x) == (x << 256) % P, hex(x)
    for nx in (1, 2, 3):
        top = (1 << (64 * nx)) - 1
        for x in [0, 1, 2, top, top - 1, 1 << (64 * nx - 1)] + [rng.getrandbits(rng.choice([1, 15, 33, 64 * nx - 1, 64 * nx])) for _ in range(count)]:
            x &= top
            assert redc_short(x, nx) == (x << 256) % P, (nx, hex(x))
    return len(tests)
