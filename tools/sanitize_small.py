"""Smallest run that touches every kernel family (for compute-sanitizer): 128-bit key (Cfg<4,19>), a few dozen units."""
import sys, random
sys.path.insert(0, '/root/repo')
from paillier_halo2_b200 import PaillierKey
from oracle.paillier_oracle import paillier_enc_native
rng = random.Random(1)
nb = 128
n = rng.getrandbits(nb) | (1 << (nb - 1)) | 1
g = rng.getrandbits(nb)
ms = [rng.getrandbits(nb) for _ in range(40)]; rs = [rng.getrandbits(nb) for _ in range(40)]
ms[0] = 0; rs[1] = 1
with PaillierKey(n, g, nb, 64) as key:
    cs = key.paillier_enc_native(ms, rs)
    assert cs[:3] == [paillier_enc_native(n, g, m, r) for m, r in zip(ms[:3], rs[:3])]
    cw, dig = key.encrypt_witness_digest(ms, rs)
    assert cw == cs
    res, q = key.paillier_add_native(cs[:20], cs[20:], want_q=True)
    t = key.tally(cs)
    cells = key.mulmod_cells([(cs[0], cs[1], q[0], res[0])][:0] + [(cs[i], cs[20 + i], q[i], res[i]) for i in range(8)], 15, montgomery=True)
    c1, allcells = key.encrypt_cells(ms[2], rs[2], 15)
print("sanitize_small ok", len(allcells))
