"""The C++ host mirror of the reference's Rust API (include/paillier_chip_host.hpp) and its test program
(tests/host/test_paillier_chip.cpp: the reference's own tests restated, src/paillier.rs:107-260)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host", "test_paillier_chip.cpp")
EXE = os.path.join(ROOT, "tests", "host", "test_paillier_chip")
LIBDIR = os.path.join(ROOT, "paillier_halo2_b200")


def _build(built_lib):
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-L" + LIBDIR, "-lpaillier_b200", "-lcrypto",
           "-Wl,-rpath," + LIBDIR, "-o", EXE]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def _env():
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = LIBDIR + ":/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    return env


def test_host_layer_builds_and_fails_loudly_without_gpu(built_lib):
    exe = _build(built_lib)
    if built_lib.pb200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, "expect-no-device"], capture_output=True, text=True, env=_env(), timeout=120)
    assert r.returncode == 0 and "CUDA failure" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_tests_in_cpp_on_gpu(built_lib):
    """test_paillier_encryption (128/64) and test_encryption_addition (264/88) of the reference, through PaillierChip::encrypt/add
    of the C++ host layer: GPU-produced witnesses and cells, constraint re-check, tamper rejection, error behaviour."""
    exe = _build(built_lib)
    r = subprocess.run([exe], capture_output=True, text=True, env=_env(), timeout=600)
    assert r.returncode == 0 and "all host tests passed" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
