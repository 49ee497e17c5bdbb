"""block28u (tcgen05.mma + TMEM for the constant-operand Barrett phases) against block28t (mma.sync) and the oracle.

The two engines are specified to leave IDENTICAL lazy digits after every modular multiplication, so the comparison is digit for digit
(pb200_debug_mulmod): the 2L-digit product of phase A (whose stash lives in TMEM for block28u), the packed quotient-estimate rows of
phase B, the value after phase C, and chains of multiplications.  Then whole batches through the public entry points."""
import ctypes as C

import numpy as np
import pytest

from oracle import cpu_ref
from oracle.paillier_oracle import tally_native
from paillier_halo2_b200 import _lib, workload
from paillier_halo2_b200.api import PaillierKey, Pb200Error, words_to_ints
from tools.umma_debug import image

pytestmark = pytest.mark.gpu


def _key_u(n_bits=2048, g="g_rand", engine=4):
    kd = workload.load_key(n_bits)
    key = PaillierKey(kd["n"], kd[g], n_bits)
    try:
        key.set_engine(engine)
    except Pb200Error as e:
        key.close()
        if e.status == _lib.PB200_ERR_UNSUPPORTED:
            pytest.skip("no block28u variant for this key size")
        raise
    return key, kd


def _lazy_values(rng, L, edge):
    beta = 14 * (2 * L - 1)
    out = []
    for lane in range(32):
        v = int.from_bytes(rng.bytes((beta - 2 + 7) // 8), "little") >> ((8 - (beta - 2) % 8) % 8)
        if lane % 3 == 1:
            v = -v
        if edge:
            v = [0, 1, -1, (1 << (beta - 2)) - 1, -((1 << (beta - 2)) - 1), (1 << 28) - 1, 1 << 27, -(1 << 27)][lane % 8] if lane < 16 else v
        out.append(v)
    return out


@pytest.mark.parametrize("engine", [4, 5])
@pytest.mark.parametrize("mode", ["sqr", "mul"])
@pytest.mark.parametrize("edge", [False, True])
def test_block28u_digits_equal_block28t(built_lib, mode, edge, engine):
    key, _ = _key_u(engine=engine)
    with key:
        lib = key._lib
        g, bl = C.c_int(), C.c_int()
        assert lib.pb200_key_shape(key.handle, C.byref(g), C.byref(bl)) == 0
        G, BL = g.value, bl.value
        L, CH = G * BL, (BL + 3) // 4
        rng = np.random.default_rng(11 + edge)
        v = image(_lazy_values(rng, L, edge), G, BL)
        y = image(_lazy_values(rng, L, edge), G, BL) if mode == "mul" else None
        yp = y.ctypes.data if y is not None else None
        outs = {}
        for eng in (3, engine):
            t = np.zeros((2 * G, CH, 32, 4), dtype=np.int32)
            assert lib.pb200_debug_mulmod(key.handle, eng, v.ctypes.data, yp, 1, None, t.ctypes.data, None) == 0
            vo = np.zeros_like(v)
            rows = np.zeros((32, L), dtype=np.uint32)
            assert lib.pb200_debug_mulmod(key.handle, eng, v.ctypes.data, yp, 1, vo.ctypes.data, None, rows.ctypes.data) == 0
            v9 = np.zeros_like(v)
            assert lib.pb200_debug_mulmod(key.handle, eng, v.ctypes.data, yp, 9, v9.ctypes.data, None, None) == 0
            outs[eng] = (t, vo, rows, v9)
        for a, b in zip(outs[3], outs[engine]):
            assert np.array_equal(a, b)
        # the product itself against Python ints (every lane)
        def val(img, lane, blocks):
            return sum(int(img[p // BL, (p % BL) // 4, lane, (p % BL) % 4]) << (28 * p) for p in range(blocks * BL))
        for lane in range(32):
            a = val(v, lane, G)
            b = val(y, lane, G) if y is not None else a
            assert val(outs[engine][0], lane, 2 * G) == a * b
        # and the lazy result is congruent to it modulo n^2 (the modulus of the engine is a multiple of n^2)
        n2 = key.n * key.n
        for lane in range(0, 32, 5):
            a = val(v, lane, G)
            b = val(y, lane, G) if y is not None else a
            assert (val(outs[engine][1], lane, G) - a * b) % n2 == 0


@pytest.mark.parametrize("engine", [4, 5])
@pytest.mark.parametrize("g", ["g_rand", "g_std"])
def test_block28u_batch_vs_cpu_and_block28t(built_lib, g, engine):
    """ragged batch (not a multiple of 32 or 64), every unit against the OpenSSL port of the oracle and against block28t"""
    count = 2500 + 33 * (engine == 5)
    key, kd = _key_u(2048, g, engine)
    with key:
        assert key.engine.startswith("block28u")
        m_w, r_w = workload.units(2048, count, seed_offset=5)
        m_w[0, :] = 0; r_w[1, :] = 0; r_w[1, 0] = 1; m_w[2, :] = np.uint64(0xFFFFFFFFFFFFFFFF); m_w[2, -1] = np.uint64((1 << 63) - 1)
        got = key.encrypt_words(m_w, r_w)
        again = key.encrypt_words(m_w, r_w)
        assert np.array_equal(got, again)          # deterministic across launches (mbarrier / TMEM reuse)
        key.set_engine(3)
        assert key.engine.startswith("block28t")
        assert np.array_equal(got, key.encrypt_words(m_w, r_w))
        want = cpu_ref.enc_batch(kd["n"], kd[g], 2048 // 64, m_w, r_w, threads=cpu_ref.hardware_threads())
        assert np.array_equal(got, want)


@pytest.mark.parametrize("engine", [4, 5])
@pytest.mark.parametrize("count", [1, 31, 33, 65, 1000, 20011])
def test_block28u_tally_and_decrypt(built_lib, count, engine):
    key, kd = _key_u(engine=engine)
    with key:
        n = kd["n"]
        cs_w = workload.ciphertexts(2048, count, n)
        got = words_to_ints(key.tally_words(cs_w).reshape(1, -1))[0]
        assert got == tally_native(n, words_to_ints(cs_w))
    key, kd = _key_u(2048, "g_std", engine)
    with key:
        p, q, n = kd["p"], kd["q"], kd["n"]
        lam = (p - 1) * (q - 1)
        key.set_private(lam, pow(lam, -1, n))
        k = min(count, 200)
        m_w, r_w = workload.units(2048, k, seed_offset=9)
        c_w = key.encrypt_words(m_w, r_w)
        assert np.array_equal(key.decrypt_words(c_w), m_w)
