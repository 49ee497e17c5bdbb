"""One-off randomized cross-checks on a B200 (not part of the test suite): block28w vs simple64 witness digests at the large key
sizes, k_add_w vs Python divmod on unreduced random pairs, fast chain vs witness chain ciphertexts on a full 2^16 batch."""
import sys, time, json, random
sys.path.insert(0, '/root/repo')
import numpy as np
from paillier_halo2_b200 import PaillierKey, workload
from paillier_halo2_b200.api import words_to_ints, ints_to_words

res = {}
for n_bits, count in ((3072, 400), (4096, 200), (2048, 1500), (1024, 9000)):
    kd = workload.load_key(n_bits)
    m_w, r_w = workload.units(n_bits, count, seed_offset=4242)
    ms, rs = words_to_ints(m_w), words_to_ints(r_w)
    with PaillierKey(kd["n"], kd["g_rand"], n_bits, 64) as key:
        t0 = time.time(); cw, dw = key.encrypt_witness_digest(ms, rs); t1 = time.time()
        key.set_engine(1)
        cs, ds = key.encrypt_witness_digest(ms, rs); t2 = time.time()
    ok = (cw == cs) and (dw == ds)
    res[f"witness_{n_bits}"] = {"units": count, "agree": ok, "block28w_s": t1 - t0, "simple64_s": t2 - t1}
    print(n_bits, count, ok, round(t1 - t0, 2), round(t2 - t1, 2), flush=True)
    assert ok
rng = random.Random(5)
for n_bits in (1024, 2048, 4096):
    n = workload.load_key(n_bits)["n"]; n2 = n * n
    N = 20000
    a = [rng.getrandbits(2 * n_bits) >> rng.choice((0, 0, 1, 5, 64)) for _ in range(N)]
    b = [rng.randrange(n2) >> rng.choice((0, 0, 3, 64, 2000)) for _ in range(N)]      # b < n^2 keeps q < a < 2^(2 n_bits)
    with PaillierKey(n, n + 1, n_bits, 64) as key:
        r_, q_ = key.paillier_add_native(a, b, want_q=True)
    bad = sum(1 for x, y, q, r in zip(a, b, q_, r_) if divmod(x * y, n2) != (q, r))
    res[f"add_{n_bits}"] = {"pairs": N, "mismatches": bad}
    print("add", n_bits, N, "mismatches", bad, flush=True)
    assert bad == 0
json.dump(res, open("/root/repo/gpurun_out/fuzz_gpu.json", "w"), indent=1)
print("fuzz ok")
