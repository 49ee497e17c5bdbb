"""Generate paillier_halo2_b200/data/keys.json (SURVEY.md §8d key recipe).  Run once; output committed."""
import json, os, sys
import numpy as np
from sympy import nextprime

out = {}
for n_bits in (128, 256, 1024, 2048, 3072, 4096):
    rng = np.random.Generator(np.random.Philox(key=20261018 + n_bits))
    half = n_bits // 2
    def draw():
        x = int.from_bytes(rng.bytes(half // 8), "little")
        x |= (3 << (half - 2))
        return x
    p = nextprime(draw()); q = nextprime(draw())
    while q == p: q = nextprime(q)
    n = p * q
    assert n.bit_length() == n_bits and n % 2 == 1
    g = int.from_bytes(rng.bytes(n_bits // 8), "little") | 2
    out[str(n_bits)] = {"p": hex(p), "q": hex(q), "n": hex(n), "g_rand": hex(g)}
    print(n_bits, "ok", file=sys.stderr)
path = os.path.join(os.path.dirname(__file__), "..", "paillier_halo2_b200", "data", "keys.json")
json.dump(out, open(path, "w"), indent=1)
