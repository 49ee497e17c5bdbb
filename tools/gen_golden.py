"""Generate tests/golden/kat.json from the CPU oracle (oracle/paillier_oracle.py).

The reference holds no golden vectors or KATs for this path and cannot be built
here, so these fixtures are produced by the Python restatement and cross-checked at generation time
against two independent exact-integer implementations: OpenSSL BIGNUM (oracle/paillier_cpu.cpp) and
GMP 6.3 (mpz_powm through ctypes).  Re-run: `python tools/gen_golden.py`.
"""
import ctypes, ctypes.util, json, os, random, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
from oracle import cpu_ref
from oracle.paillier_oracle import (paillier_enc_native, paillier_add_native, tally_native, encrypt_steps,
                                    pow_chain_steps)
from paillier_halo2_b200 import workload
from paillier_halo2_b200.api import witness_digest, ints_to_words, words_to_ints

gmp = ctypes.CDLL("libgmp.so.10")


class MPZ(ctypes.Structure):
    _fields_ = [("alloc", ctypes.c_int), ("size", ctypes.c_int), ("d", ctypes.c_void_p)]


def gmp_powm(b, e, m):
    xs = [MPZ() for _ in range(4)]
    for x, v in zip(xs[1:], (b, e, m)):
        gmp.__gmpz_init_set_str(ctypes.byref(x), hex(v)[2:].encode(), 16)
    gmp.__gmpz_init(ctypes.byref(xs[0]))
    gmp.__gmpz_powm(ctypes.byref(xs[0]), ctypes.byref(xs[1]), ctypes.byref(xs[2]), ctypes.byref(xs[3]))
    buf = ctypes.create_string_buffer(gmp.__gmpz_sizeinbase(ctypes.byref(xs[0]), 16) + 2)
    gmp.__gmpz_get_str(buf, 16, ctypes.byref(xs[0]))
    for x in xs:
        gmp.__gmpz_clear(ctypes.byref(x))
    return int(buf.value, 16)


def gmp_enc(n, g, m, r):
    n2 = n * n
    return gmp_powm(g, m, n2) * gmp_powm(r, n, n2) % n2


rng = random.Random(0x5041494C)
H = lambda v: hex(v)
out = {"enc": [], "add": [], "tally": [], "witness": [], "errors": []}


def enc_case(n_bits, limb_bits, n, g, m, r, tag):
    c = paillier_enc_native(n, g, m, r)
    assert c == gmp_enc(n, g, m, r), "GMP disagrees"
    if n_bits % 64 == 0:
        w = n_bits // 64
        o = cpu_ref.enc_batch(n, g, w, ints_to_words([m], w), ints_to_words([r], w))
        assert words_to_ints(o)[0] == c, "OpenSSL disagrees"
    out["enc"].append({"tag": tag, "n_bits": n_bits, "limb_bits": limb_bits, "n": H(n), "g": H(g), "m": H(m), "r": H(r), "c": H(c)})


# reference default sizes (src/paillier.rs:115-116, :186-187), odd n drawn like rng.gen_biguint
for n_bits, limb_bits in ((128, 64), (264, 88)):
    for t in range(4):
        n = rng.getrandbits(n_bits) | 1
        enc_case(n_bits, limb_bits, n, rng.getrandbits(n_bits), rng.getrandbits(n_bits), rng.getrandbits(n_bits), f"ref-default-{t}")
    n = rng.getrandbits(n_bits) | 1
    full = (1 << n_bits) - 1
    for tag, g, m, r in (("m=0", rng.getrandbits(n_bits), 0, rng.getrandbits(n_bits)), ("m=1", rng.getrandbits(n_bits), 1, rng.getrandbits(n_bits)),
                         ("r=1", rng.getrandbits(n_bits), rng.getrandbits(n_bits), 1), ("r=0", rng.getrandbits(n_bits), rng.getrandbits(n_bits), 0),
                         ("g=0", 0, rng.getrandbits(n_bits), rng.getrandbits(n_bits)), ("all-ones", full, full, full),
                         ("m=2^k", rng.getrandbits(n_bits), 1 << (n_bits - 1), rng.getrandbits(n_bits)),
                         ("small-n", 5, 7, 9)):
        nn = 3 if tag == "small-n" else n
        enc_case(n_bits, limb_bits, nn, g, m, r, tag)
    enc_case(n_bits, limb_bits, 1, rng.getrandbits(n_bits), rng.getrandbits(n_bits), rng.getrandbits(n_bits), "n=1")
    enc_case(n_bits, limb_bits, full, full, full, full, "n=all-ones")

# production sizes with the seeded Paillier keys
for n_bits in (256, 1024, 2048, 3072, 4096):
    key = workload.load_key(n_bits)
    n = key["n"]
    m_w, r_w = workload.units(n_bits, 3)
    ms, rs = words_to_ints(m_w), words_to_ints(r_w)
    for i in range(2 if n_bits > 2048 else 3):
        enc_case(n_bits, 64, n, key["g_rand"], ms[i], rs[i], f"philox-unit-{i}-g_rand")
    enc_case(n_bits, 64, n, n + 1, ms[0], rs[0], "philox-unit-0-g=n+1")
    assert paillier_enc_native(n, n + 1, ms[0], 1) == (1 + ms[0] * n) % (n * n)  # README.md:10 identity
    if n_bits <= 2048:
        enc_case(n_bits, 64, n, key["g_rand"], 0, rs[1], "m=0")
        enc_case(n_bits, 64, n, key["g_rand"], n - 1, rs[1], "m=n-1")
        enc_case(n_bits, 64, n, key["g_rand"], ms[1], 1, "r=1")
        enc_case(n_bits, 64, n, key["g_rand"], 1 << (n_bits - 1), n - 1, "m=2^k,r=n-1")
        enc_case(n_bits, 64, n, (1 << n_bits) - 1, (1 << n_bits) - 1, (1 << n_bits) - 1, "all-ones")

# add: half-width inputs like the reference's test (src/paillier.rs:216-221) and real ciphertexts
for n_bits, limb_bits in ((128, 64), (264, 88), (1024, 64), (2048, 64)):
    n = (rng.getrandbits(n_bits) | 1) if n_bits < 1024 else workload.load_key(n_bits)["n"]
    n2 = n * n
    for t in range(3):
        c1, c2 = rng.getrandbits(n_bits), rng.getrandbits(n_bits)
        out["add"].append({"n_bits": n_bits, "limb_bits": limb_bits, "n": H(n), "c_bits": n_bits, "c1": H(c1), "c2": H(c2),
                           "res": H(paillier_add_native(n, c1, c2)), "q": H(c1 * c2 // n2)})
    for c1, c2 in ((rng.randrange(n2), rng.randrange(n2)), (n2 - 1, n2 - 1), (0, n2 - 1), (1, 1)):
        out["add"].append({"n_bits": n_bits, "limb_bits": limb_bits, "n": H(n), "c_bits": 2 * n_bits, "c1": H(c1), "c2": H(c2),
                           "res": H(paillier_add_native(n, c1, c2)), "q": H(c1 * c2 // n2)})
    cs = [rng.randrange(n2) for _ in range(37)]
    out["tally"].append({"n_bits": n_bits, "n": H(n), "cs": [H(c) for c in cs], "res": H(tally_native(n, cs))})

# witness: full (q, rem) streams at the reference default sizes, digests at production sizes
for n_bits, limb_bits, full_records in ((128, 64, True), (264, 88, True), (1024, 64, False), (2048, 64, False)):
    if n_bits < 1024:
        n, g = rng.getrandbits(n_bits) | 1, rng.getrandbits(n_bits)
        units = [(rng.getrandbits(n_bits), rng.getrandbits(n_bits)) for _ in range(2)] + [(0, 5), (1, 1), ((1 << n_bits) - 1, (1 << n_bits) - 1)]
    else:
        key = workload.load_key(n_bits)
        n, g = key["n"], key["g_rand"]
        m_w, r_w = workload.units(n_bits, 2)
        units = list(zip(words_to_ints(m_w), words_to_ints(r_w)))
    wo = (2 * n_bits + 63) // 64
    n2 = n * n
    _, gsteps = pow_chain_steps(g, (1 << n_bits) - 1, n2)
    gsq = [s for s in gsteps if s.kind == "sqr"]
    gchain_digest = witness_digest([(s.q, s.rem) for s in gsq], wo)
    case = {"n_bits": n_bits, "limb_bits": limb_bits, "n": H(n), "g": H(g), "g_chain_digest": H(gchain_digest), "units": []}
    for m, r in units:
        c, steps = encrypt_steps(n, g, m, r)
        mb = m.bit_length()
        gpart = steps[: mb + bin(m).count("1")]
        per_unit = [s for s in gpart if s.kind == "mul"] + steps[len(gpart):]
        u = {"m": H(m), "r": H(r), "c": H(c), "n_records": len(per_unit), "g_mul_count": bin(m).count("1"),
             "digest": H(witness_digest([(s.q, s.rem) for s in per_unit], wo))}
        if full_records and (n_bits == 128 or len(case["units"]) < 1 or m < 2):
            u["records"] = [[H(s.q), H(s.rem)] for s in per_unit]
        else:
            u["first"] = [H(per_unit[0].q), H(per_unit[0].rem)]
            u["last"] = [H(per_unit[-1].q), H(per_unit[-1].rem)]
        case["units"].append(u)
    out["witness"].append(case)

path = os.path.join(ROOT, "tests", "golden", "kat.json")
json.dump(out, open(path, "w"), indent=0)
print(path, os.path.getsize(path), {k: len(v) for k, v in out.items()})
