"""Deterministic synthetic inputs for the Paillier hot path.

Keys: n = p*q with |n|/2-bit primes (top two bits set, so n has exactly |n| bits and is odd),
generated once by tools/gen_keys.py with sympy.nextprime from a Philox(key = 20261018 + |n|) stream
and committed as data/keys.json.  g = n + 1 (standard) and a second, random g in [2, 2^|n|).
Units: m_i, r_i uniform in [0, 2^(|n|-1)) — the top bit of the top word is cleared, so they are
< n without a bignum reduction — from one Philox(key = 0x5041494C4C494552) stream in unit order
(prefix-stable: the first N units of a larger batch are the same values).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple

import numpy as np

UNIT_KEY = 0x5041494C4C494552
_KEYS: Dict[int, dict] = {}


def load_key(n_bits: int) -> dict:
    """{'n': int, 'g_std': n+1, 'g_rand': int, 'p': int, 'q': int} for |n| = n_bits."""
    if not _KEYS:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "keys.json")
        with open(path) as f:
            raw = json.load(f)
        for k, v in raw.items():
            _KEYS[int(k)] = {kk: int(vv, 16) for kk, vv in v.items()}
    key = dict(_KEYS[n_bits])
    key["g_std"] = key["n"] + 1
    return key


def units(n_bits: int, count: int, seed_offset: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(m, r): two (count, n_bits/64) uint64 little-endian arrays."""
    assert n_bits % 64 == 0
    w = n_bits // 64
    rng = np.random.Generator(np.random.Philox(key=UNIT_KEY + seed_offset))
    raw = rng.integers(0, 1 << 64, size=(count, 2, w), dtype=np.uint64, endpoint=False)
    raw[:, :, w - 1] &= np.uint64((1 << 63) - 1)
    return np.ascontiguousarray(raw[:, 0, :]), np.ascontiguousarray(raw[:, 1, :])


def ciphertexts(n_bits: int, count: int, n: int, seed_offset: int = 1) -> np.ndarray:
    """(count, 2*n_bits/64) uniform values below n^2 (top bits cleared below bitlen(n^2)-1) for the tally."""
    w = 2 * n_bits // 64
    rng = np.random.Generator(np.random.Philox(key=UNIT_KEY + seed_offset))
    raw = rng.integers(0, 1 << 64, size=(count, w), dtype=np.uint64, endpoint=False)
    top_bits = (n * n).bit_length() - 1  # values < 2^top_bits <= n^2
    for i in range(w):
        lo = 64 * i
        if lo >= top_bits:
            raw[:, i] = 0
        elif top_bits - lo < 64:
            raw[:, i] &= np.uint64((1 << (top_bits - lo)) - 1)
    return raw
