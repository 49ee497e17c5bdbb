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
