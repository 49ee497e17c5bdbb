"""ctypes wrapper of oracle/_build/libpaillier_cpu.so (OpenSSL BIGNUM port of src/paillier.rs:87-97).

TEST / BASELINE INFRASTRUCTURE ONLY (tests/, bench.py cpu_baseline and --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libpaillier_cpu.so")
u64p = C.POINTER(C.c_uint64)
_lib = None


def build() -> str:
    subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        lib = C.CDLL(LIB_PATH)
        lib.cpu_hardware_threads.restype = C.c_int
        lib.cpu_paillier_enc_batch.restype = C.c_int
        lib.cpu_paillier_enc_batch.argtypes = [u64p, u64p, C.c_int, u64p, u64p, C.c_size_t, u64p, C.c_int]
        lib.cpu_paillier_add_batch.restype = C.c_int
        lib.cpu_paillier_add_batch.argtypes = [u64p, C.c_int, u64p, u64p, C.c_int, C.c_size_t, u64p, C.c_int]
        lib.cpu_paillier_tally.restype = C.c_int
        lib.cpu_paillier_tally.argtypes = [u64p, C.c_int, u64p, C.c_size_t, u64p, C.c_int]
        lib.cpu_witness_digest_batch.restype = C.c_int
        lib.cpu_witness_digest_batch.argtypes = [u64p, u64p, C.c_int, u64p, u64p, C.c_size_t, u64p, u64p, C.c_int, C.c_int]
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(u64p)


def _w(v: int, n: int) -> np.ndarray:
    return np.frombuffer(int(v).to_bytes(8 * n, "little"), dtype="<u8").copy()


def hardware_threads() -> int:
    return load().cpu_hardware_threads()


def enc_batch(n: int, g: int, words_in: int, m_w: np.ndarray, r_w: np.ndarray, threads: int = 1) -> np.ndarray:
    lib = load()
    m_w = np.ascontiguousarray(m_w, dtype="<u8")
    r_w = np.ascontiguousarray(r_w, dtype="<u8")
    count = m_w.shape[0]
    out = np.empty((count, 2 * words_in), dtype="<u8")
    rc = lib.cpu_paillier_enc_batch(_p(_w(n, words_in)), _p(_w(g, words_in)), words_in, _p(m_w), _p(r_w), count, _p(out), threads)
    assert rc == 0
    return out


def add_batch(n: int, words_in: int, c1_w: np.ndarray, c2_w: np.ndarray, threads: int = 1) -> np.ndarray:
    lib = load()
    c1_w = np.ascontiguousarray(c1_w, dtype="<u8")
    c2_w = np.ascontiguousarray(c2_w, dtype="<u8")
    count, cw = c1_w.shape
    out = np.empty((count, 2 * words_in), dtype="<u8")
    lib.cpu_paillier_add_batch(_p(_w(n, words_in)), words_in, _p(c1_w), _p(c2_w), cw, count, _p(out), threads)
    return out


def tally(n: int, words_in: int, c_w: np.ndarray, threads: int = 1) -> np.ndarray:
    lib = load()
    c_w = np.ascontiguousarray(c_w, dtype="<u8").reshape(-1, 2 * words_in)
    out = np.empty(2 * words_in, dtype="<u8")
    lib.cpu_paillier_tally(_p(_w(n, words_in)), words_in, _p(c_w), c_w.shape[0], _p(out), threads)
    return out



def witness_digest_batch(n: int, g: int, words_in: int, m_w: np.ndarray, r_w: np.ndarray, threads: int = 1, backend: str = "auto"):
    """(ciphertexts, digests, backend used) of the reference's mul_mod chain per unit (src/paillier.rs:51,55,57; SURVEY.md
    A.4-A.5): full product + div_rem per step, digest as defined in include/paillier_b200.h.  backend: "openssl" (BN_div),
    "gmp" (mpz_tdiv_qr, 2.5x faster) or "auto" (gmp when libgmp.so.10 loads).  Raises if some quotient overflows words_out."""
    lib = load()
    m_w = np.ascontiguousarray(m_w, dtype="<u8")
    r_w = np.ascontiguousarray(r_w, dtype="<u8")
    count = m_w.shape[0]
    c = np.empty((count, 2 * words_in), dtype="<u8")
    dig = np.empty(count, dtype="<u8")
    args = (_p(_w(n, words_in)), _p(_w(g, words_in)), words_in, _p(m_w), _p(r_w), count, _p(c), _p(dig), threads)
    rc, used = -1, "openssl"
    if backend in ("auto", "gmp"):
        rc, used = lib.cpu_witness_digest_batch(*args, 1), "gmp"
        if rc == -1 and backend == "gmp":
            raise RuntimeError("libgmp.so.10 not loadable")
    if rc == -1:
        rc, used = lib.cpu_witness_digest_batch(*args, 0), "openssl"
    if rc != 0:
        raise OverflowError("a mul_mod quotient does not fit 2*enc_bits (range check on q fails)")
    return c, dig, used
