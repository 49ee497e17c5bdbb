"""CPU tests of the oracle: against the committed golden vectors, against OpenSSL BIGNUM, against
Paillier's own algebra (decryption with the fixture primes), and the reference's test flows
(src/paillier.rs:113-259) through the chip restatement + constraint re-checker."""
import random

import numpy as np
import pytest

from oracle import cpu_ref
from oracle.paillier_oracle import (BigUintChip, Context, RefreshAux, check_constraints, decompose, encrypt_steps,
                                    get_biguint, paillier_add_native, paillier_enc_add_test, paillier_enc_native,
                                    paillier_enc_test, pow_chain_steps, tally_native)
from paillier_halo2_b200 import workload
from paillier_halo2_b200.api import ints_to_words, witness_digest, words_to_ints
from util import h, kat


def test_golden_enc_vectors():
    for case in kat()["enc"]:
        assert paillier_enc_native(h(case["n"]), h(case["g"]), h(case["m"]), h(case["r"])) == h(case["c"]), case["tag"]


def test_golden_add_and_tally_vectors():
    for case in kat()["add"]:
        n = h(case["n"])
        assert paillier_add_native(n, h(case["c1"]), h(case["c2"])) == h(case["res"])
        assert h(case["c1"]) * h(case["c2"]) // (n * n) == h(case["q"])
    for case in kat()["tally"]:
        assert tally_native(h(case["n"]), [h(c) for c in case["cs"]]) == h(case["res"])


def test_golden_witness_streams():
    for case in kat()["witness"]:
        n, g = h(case["n"]), h(case["g"])
        wo = (2 * case["n_bits"] + 63) // 64
        for u in case["units"]:
            m, r = h(u["m"]), h(u["r"])
            c, steps = encrypt_steps(n, g, m, r)
            assert c == h(u["c"])
            ng = m.bit_length() + bin(m).count("1")
            per_unit = [s for s in steps[:ng] if s.kind == "mul"] + steps[ng:]
            assert len(per_unit) == u["n_records"]
            assert witness_digest([(s.q, s.rem) for s in per_unit], wo) == h(u["digest"])
            if "records" in u:
                assert [[hex(s.q), hex(s.rem)] for s in per_unit] == u["records"]
            for s in steps:  # q, rem are THE quotient and remainder
                assert s.a * s.b == s.q * n * n + s.rem and 0 <= s.rem < n * n


@pytest.mark.parametrize("n_bits", [1024, 2048])
def test_oracle_matches_openssl(n_bits):
    key = workload.load_key(n_bits)
    w = n_bits // 64
    m_w, r_w = workload.units(n_bits, 6)
    got = words_to_ints(cpu_ref.enc_batch(key["n"], key["g_rand"], w, m_w, r_w, threads=2))
    ms, rs = words_to_ints(m_w), words_to_ints(r_w)
    assert got == [paillier_enc_native(key["n"], key["g_rand"], m, r) for m, r in zip(ms, rs)]
    c_w = workload.ciphertexts(n_bits, 50, key["n"])
    cs = words_to_ints(c_w)
    assert words_to_ints(cpu_ref.tally(key["n"], w, c_w, threads=3))[0] == tally_native(key["n"], cs)
    s_w = cpu_ref.add_batch(key["n"], w, c_w[:25], c_w[25:], threads=2)
    assert words_to_ints(s_w) == [paillier_add_native(key["n"], a, b) for a, b in zip(cs[:25], cs[25:])]


@pytest.mark.parametrize("n_bits", [256, 1024])
def test_paillier_algebra_roundtrip(n_bits):
    """encode -> (homomorphic add) -> decode with the fixture primes: D(E(m)) = m, D(E(a)E(b)) = a+b mod n."""
    key = workload.load_key(n_bits)
    n, p, q = key["n"], key["p"], key["q"]
    n2 = n * n
    lam = (p - 1) * (q - 1)
    g = n + 1
    mu = pow((pow(g, lam, n2) - 1) // n, -1, n)
    dec = lambda c: (pow(c, lam, n2) - 1) // n * mu % n
    m_w, r_w = workload.units(n_bits, 4)
    ms, rs = words_to_ints(m_w), words_to_ints(r_w)
    cs = [paillier_enc_native(n, g, m, r) for m, r in zip(ms, rs)]
    assert [dec(c) for c in cs] == ms
    assert dec(paillier_add_native(n, cs[0], cs[1])) == (ms[0] + ms[1]) % n
    assert dec(tally_native(n, cs)) == sum(ms) % n
    assert paillier_enc_native(n, g, ms[0], 1) == (1 + ms[0] * n) % n2  # README.md:10


def test_chain_is_modpow():
    rng = random.Random(7)
    for _ in range(20):
        n2 = (rng.getrandbits(200) | 1) ** 2
        a, e = rng.getrandbits(256), rng.getrandbits(rng.choice([0, 1, 5, 64, 130]))
        acc, steps = pow_chain_steps(a, e, n2)
        assert acc == pow(a, e, n2)
        assert len(steps) == e.bit_length() + bin(e).count("1")


@pytest.mark.parametrize("enc_bits,limb_bits", [(128, 64), (264, 88)])
def test_reference_test_flows(enc_bits, limb_bits):
    """src/paillier.rs:113-182 and :184-259 at the repo's default sizes, k=16 / lookup_bits=15."""
    rng = random.Random(enc_bits)
    n = rng.getrandbits(enc_bits) | 1
    g, m, r = (rng.getrandbits(enc_bits) for _ in range(3))
    ctx = paillier_enc_test(enc_bits, limb_bits, n, g, m, r, paillier_enc_native(n, g, m, r), lookup_bits=15)
    assert len(ctx.steps) == m.bit_length() + bin(m).count("1") + n.bit_length() + bin(n).count("1") + 1
    c1, c2 = rng.getrandbits(enc_bits), rng.getrandbits(enc_bits)
    ctx = paillier_enc_add_test(enc_bits, limb_bits, n, g, c1, c2, paillier_add_native(n, c1, c2), lookup_bits=15)
    assert len(ctx.steps) == 1
    with pytest.raises(AssertionError):  # wrong expected result is rejected like assert_eq!/assert_equal_fresh
        paillier_enc_add_test(enc_bits, limb_bits, n, g, c1, c2, paillier_add_native(n, c1, c2) ^ 1)


def test_constraint_checker_rejects_bad_witness():
    big = BigUintChip(64)
    ctx = Context()
    n = big.assign_integer(ctx, (1 << 127) + 12345, 128)
    a = big.assign_integer(ctx, 3 << 100, 128)
    big.mul_mod(ctx, a, a, n)
    check_constraints(ctx)
    ctx.checks.append(("tampered", False))
    with pytest.raises(AssertionError):
        check_constraints(ctx)
    with pytest.raises(AssertionError):
        big.assign_integer(Context(), 5, 100)  # bit_len % limb_bits != 0


def test_refresh_aux_sizes_and_limb_order():
    # Fresh n^2 limb counts
    for lb, nl, want in ((64, 2, 4), (88, 3, 6), (64, 16, 32), (64, 32, 64), (64, 48, 96), (64, 64, 128)):
        aux = RefreshAux(lb, nl, nl)
        assert aux.num_limbs_out == want and max(aux.increased_limbs_vec) == 2
    v = 0x1234567890ABCDEF_FEDCBA0987654321_0F1E2D3C4B5A6978
    assert get_biguint(decompose(v, 3, 88), 88) == v and decompose(v, 3, 64)[0] == 0x0F1E2D3C4B5A6978


def test_zero_modulus_panics_like_num_bigint():
    with pytest.raises(ZeroDivisionError):
        paillier_enc_native(0, 1, 2, 3)
    with pytest.raises(ZeroDivisionError):
        paillier_add_native(0, 1, 2)


def _cell_hash(cells):
    import hashlib
    h = hashlib.sha256()
    for c in cells:
        h.update(int(c).to_bytes(32, "little"))
    return h.hexdigest()


def test_golden_cell_streams():
    """tests/golden/cells.json (tools/gen_golden_cells.py): the chip restatement reproduces the committed cell streams."""
    import json, os
    from oracle.paillier_oracle import Assigned, BigUintChip, Context, decompose, paillier_enc_test
    from paillier_halo2_b200 import workload
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cells.json")))
    for f in gold["flows"]:
        hx = lambda k: int(f[k], 16)
        ctx = paillier_enc_test(f["enc_bits"], f["limb_bits"], hx("n"), hx("g"), hx("m"), hx("r"), hx("c"), lookup_bits=f["lookup_bits"])
        assert len(ctx.cells) == f["n_cells"] and _cell_hash(ctx.cells) == f["sha256"]
        assert [hex(v) for v in ctx.cells[:6]] == f["first"] and [hex(v) for v in ctx.cells[-6:]] == f["last"]
    for gq in gold["groups"]:
        n = workload.load_key(gq["n_bits"])["n"]
        L = 2 * gq["n_bits"] // 64
        a, b = int(gq["a"], 16), int(gq["b"], 16)
        ctx = Context()
        out = BigUintChip(64, gq["lookup_bits"]).mul_mod(ctx, Assigned(decompose(a, L, 64), a, 64), Assigned(decompose(b, L, 64), b, 64),
                                                         Assigned(decompose(n * n, L, 64), n * n, 64))
        assert out.value == int(gq["rem"], 16) and len(ctx.cells) == gq["n_cells"] and _cell_hash(ctx.cells) == gq["sha256"]


@pytest.mark.parametrize("backend", ["openssl", "gmp"])
def test_cpu_witness_digest_batch_matches_python_chain(backend):
    """oracle/paillier_cpu.cpp cpu_witness_digest_batch (the full-batch checker of bench.py and the GPU tests) against the
    Python restatement of the chain, both backends: the reference's default sizes with even / short n
    (the reference draws n = gen_biguint(bits)), edge units, and a production size."""
    rng = random.Random(20261018)
    cases = []
    for n_bits in (128, 264):
        wi = (n_bits + 63) // 64
        for _ in range(3):
            n = rng.getrandbits(n_bits) | (1 << (n_bits - 1))          # full width so that q fits 2*enc_bits
            g = rng.getrandbits(n_bits)
            ms = [0, 1, (1 << n_bits) - 1] + [rng.getrandbits(n_bits) for _ in range(5)]
            rs = [1, 0, (1 << n_bits) - 1] + [rng.getrandbits(n_bits) for _ in range(5)]
            cases.append((n_bits, wi, n, g, ms, rs))
    key = workload.load_key(1024)
    m_w, r_w = workload.units(1024, 3)
    cases.append((1024, 16, key["n"], key["g_rand"], words_to_ints(m_w), words_to_ints(r_w)))
    for n_bits, wi, n, g, ms, rs in cases:
        try:
            c, dig, used = cpu_ref.witness_digest_batch(n, g, wi, ints_to_words(ms, wi), ints_to_words(rs, wi), threads=2, backend=backend)
        except RuntimeError:
            pytest.skip("libgmp.so.10 not loadable")
        assert used == backend
        for i, (m, r) in enumerate(zip(ms, rs)):
            cc, steps = encrypt_steps(n, g, m, r)
            ng = m.bit_length() + bin(m).count("1")
            mine = [(s.q, s.rem) for s in steps[:ng] if s.kind == "mul"] + [(s.q, s.rem) for s in steps[ng:]]
            assert words_to_ints(c[i:i + 1])[0] == cc
            assert int(dig[i]) == witness_digest(mine, 2 * wi)
