"""CPU tests of the boundary: the library builds, loads, exports every symbol include/paillier_b200.h
declares, validates arguments like the reference's assertions, and has NO CPU fallback."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from paillier_halo2_b200 import _lib
from paillier_halo2_b200.api import PaillierKey, Pb200Error, ints_to_words

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported_and_bound(built_lib):
    header = open(os.path.join(ROOT, "include", "paillier_b200.h")).read()
    declared = set(re.findall(r"\b(pb200_[a-z0-9_]+)\s*\(", header)) - {"pb200_witness_sink_fn"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert hasattr(built_lib, name), name


def test_strerror_and_version(built_lib):
    assert built_lib.pb200_strerror(0) == b"ok"
    for code in range(-9, 0):
        assert built_lib.pb200_strerror(code) not in (b"", b"unknown status")
    assert b"sm_100a" in built_lib.pb200_version()


def _create(lib, n, g, n_bits, limb_bits):
    win = max(1, (n_bits + 63) // 64)
    h = C.c_void_p()
    rc = lib.pb200_key_create(0, n_bits, limb_bits, ints_to_words([n], win).ctypes.data_as(_lib.u64p),
                              ints_to_words([g], win).ctypes.data_as(_lib.u64p), C.byref(h))
    if rc == 0:
        lib.pb200_key_destroy(h)
    return rc


def test_key_validation_mirrors_reference_assertions(built_lib):
    lib = built_lib
    assert _create(lib, 0, 3, 128, 64) == _lib.PB200_ERR_ZERO_MODULUS        # num-bigint panics on zero modulus
    assert _create(lib, 10, 3, 128, 64) in (_lib.PB200_OK, _lib.PB200_ERR_CUDA)      # even n is accepted, as in the reference (ERR_CUDA: no device here)
    assert _create(lib, 11, 3, 128, 60) == _lib.PB200_ERR_INVALID_ARG        # assign_integer: bit_len % limb_bits
    assert _create(lib, 11, 3, 0, 64) == _lib.PB200_ERR_INVALID_ARG
    assert _create(lib, (1 << 127) | 1, 3, 127 * 64 + 64, 64) == _lib.PB200_ERR_UNSUPPORTED  # > 4096-bit n
    assert _create(lib, (1 << 100) | 1, 1 << 90, 88, 88) == _lib.PB200_ERR_RANGE  # n does not fit 88 bits
    h = C.c_void_p()
    assert lib.pb200_key_create(0, 128, 64, None, None, C.byref(h)) == _lib.PB200_ERR_INVALID_ARG


def test_no_cpu_fallback_without_gpu(built_lib):
    """On a box without a CUDA device a valid key cannot be created: the product path fails loudly."""
    if built_lib.pb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    assert _create(built_lib, 11, 3, 128, 64) == _lib.PB200_ERR_CUDA
    with pytest.raises(Pb200Error) as e:
        PaillierKey(11, 3, 128, 64)
    assert e.value.status == _lib.PB200_ERR_CUDA


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "paillier_halo2_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(root, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_new_entry_points_reject_null_arguments(built_lib):
    """argument validation of the witness-digest, add and cell entry points needs no device"""
    lib = built_lib
    null = C.c_void_p()
    assert lib.pb200_encrypt_witness_digest_dev(null, None, None, 1, None, None) == _lib.PB200_ERR_INVALID_ARG
    assert lib.pb200_mulmod_cells_batch(null, None, None, None, None, 1, 15, 0, None) == _lib.PB200_ERR_INVALID_ARG
    assert lib.pb200_mulmod_cells_batch_dev(null, None, None, None, None, 1, 15, 0, None) == _lib.PB200_ERR_INVALID_ARG
    assert lib.pb200_assign_cells_batch(null, None, 1, 128, 15, 0, None) == _lib.PB200_ERR_INVALID_ARG
    assert lib.pb200_key_n2_cells(null, 15, 0, None) == _lib.PB200_ERR_INVALID_ARG
    assert lib.pb200_cells_layout(null, 15, None) == _lib.PB200_ERR_INVALID_ARG
    assert lib.pb200_key_witness_engine(null) == b""


def test_rust_sys_crate_declares_every_symbol():
    """rust/paillier-b200-sys cannot be compiled here (no Rust toolchain); at least its extern block must list exactly the symbols of
    the header, with the same number of parameters."""
    header = open(os.path.join(ROOT, "include", "paillier_b200.h")).read()
    rust = open(os.path.join(ROOT, "rust", "paillier-b200-sys", "src", "lib.rs")).read()
    decl_h = {}
    for mm in re.finditer(r"\b(pb200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, re.S):
        name, args = mm.group(1), mm.group(2).strip()
        if name == "pb200_witness_sink_fn":
            continue
        decl_h[name] = 0 if args in ("", "void") else args.count(",") + 1
    ext = rust[rust.index('extern "C" {'):]
    ext = ext[:ext.index("\n}\n")]
    decl_r = {}
    for mm in re.finditer(r"pub fn (pb200_[a-z0-9_]+)\s*\(([^;]*?)\)\s*(->[^;]*)?;", ext, re.S):
        args = mm.group(2).strip()
        decl_r[mm.group(1)] = 0 if not args else args.count(":")
    assert set(decl_h) == set(decl_r), set(decl_h) ^ set(decl_r)
    assert decl_h == decl_r, {k: (decl_h[k], decl_r[k]) for k in decl_h if decl_h[k] != decl_r[k]}
