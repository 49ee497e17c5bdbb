"""CPU test of the witness step's arithmetic (tests/model_block28.py: limb-exact model of block28.cuh's
mulmod_w / w_tail): q-hat from the engine's own Barrett phases + the exact tail must give
q = floor(a b / n^2), rem = a b mod n^2 and the next chain operand, for random and adversarial operands."""
import random

import pytest

import model_block28 as mb


@pytest.mark.parametrize("G,BL,n_bits,n_keys,n_rand", [(4, 19, 128, 5, 12), (4, 19, 264, 4, 8), (4, 19, 1024, 2, 2)])
def test_witness_step_exact(G, BL, n_bits, n_keys, n_rand):
    rng = random.Random(1234 + n_bits)
    P = mb.Params(G, BL)
    wo = (2 * n_bits + 63) // 64
    stats = {}
    for tr in range(n_keys):
        n = rng.getrandbits(n_bits) | (1 << (n_bits - 1)) | 1
        if tr == 1:
            n = (1 << n_bits) - 1
        if tr == 2:
            n = (1 << (n_bits - 1)) + 1
        N = n * n
        keyw, inv = mb.witness_key(P, N)
        s = keyw[0] >> 1
        assert keyw[0] % 2 == 0
        mb._cache.clear()
        top = (1 << n_bits) - 1
        cases = [(rng.randrange(N), rng.randrange(N)) for _ in range(n_rand)]
        cases += [(1, 1), (0, 5), (0, 0), (N - 1, N - 1), (1, N - 1), (top, top), (n, n), (n + 1, n - 1), (N - 1, 1), (N - n, N - n)]
        for a, b in cases:
            for sqr in (False, True):
                if sqr:
                    b = a
                A, B = mb.to_digits(a << s, P.L), mb.to_digits(b << s, P.L)
                q, rem, nxt = mb.witness_step(P, keyw, inv, A, B, wo, sqr, stats)
                assert (q, rem) == divmod(a * b, N)
                assert nxt == mb.to_digits(rem << s, P.L)
    assert stats["passes"] <= 2


def test_witness_chain_model_matches_oracle_stream():
    """The whole k_witness chain on the limb model (128-bit key, Cfg<4,19>): g-chain multiplications over the set bits of m against
    the per-key table g^(2^i) 2^s, r-chain with the multiplication of an iteration computed BEFORE its squaring but emitted after
    it, final gm*rn — the emitted (q, rem) stream must be the oracle's, record for record."""
    from oracle.paillier_oracle import encrypt_steps
    rng = random.Random(2026)
    n_bits = 128
    P = mb.Params(4, 19)
    wo = (2 * n_bits + 63) // 64
    n = rng.getrandbits(n_bits) | (1 << (n_bits - 1)) | 1
    g = rng.getrandbits(n_bits)
    N = n * n
    keyw, inv = mb.witness_key(P, N)
    s = keyw[0] >> 1
    mb._cache.clear()
    D = lambda v: mb.to_digits(v << s, P.L)
    step = lambda A, B, sqr=False: mb.witness_step(P, keyw, inv, A, B, wo, sqr)
    # per-key table (k_gchain_w): gtab[i] = g^(2^i) 2^s, with the g-chain squaring records
    gtab, cur = [], D(g)
    for _ in range(n_bits):
        gtab.append(cur)
        _, _, cur = step(cur, cur, True)
    for m, r in ((rng.getrandbits(n_bits), rng.getrandbits(n_bits)), (0, 1), ((1 << n_bits) - 1, n + 1)):
        stream = []
        acc = D(1)
        for i in range(m.bit_length()):
            if (m >> i) & 1:
                q, rem, acc = step(acc, gtab[i])
                stream.append((q, rem))
        gm = acc
        cur, acc = D(r), D(1)
        for i in range(n.bit_length()):
            pending = None
            if (n >> i) & 1:
                q, rem, acc = step(cur, acc)          # mul first: cur stays in V
                pending = (q, rem)
            q, rem, cur = step(cur, cur, True)
            stream.append((q, rem))
            if pending:
                stream.append(pending)
        q, rem, _ = step(gm, acc)
        stream.append((q, rem))
        c, steps = encrypt_steps(n, g, m, r)
        gs = m.bit_length() + bin(m).count("1")
        want = [(x.q, x.rem) for x in steps[:gs] if x.kind == "mul"] + [(x.q, x.rem) for x in steps[gs:]]
        assert rem == c and stream == want


def test_montgomery_cios32_model():
    """the 32-bit-word CIOS of the K4 Montgomery output (cells.cu: fr_to_mont32), word for word, against (x << 256) % p"""
    import model_cios32
    assert model_cios32.self_test(3000) > 3000
