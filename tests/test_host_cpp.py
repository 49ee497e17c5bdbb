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


CHECK_SRC = os.path.join(ROOT, "tests", "host", "check_cells.cpp")
CHECK_EXE = os.path.join(ROOT, "tests", "host", "check_cells")


def _oracle_flow_file(path, enc_bits, limb_bits, lookup_bits, flip=None, add=False):
    """Run the chip restatement on a seeded flow, recording where every assign_integer / square+refresh / mul_mod starts, and write
    the cells plus that bookkeeping in the format tests/host/check_cells.cpp reads."""
    import random
    from oracle import paillier_oracle as po

    rng = random.Random(99 + enc_bits + (1 if add else 0))
    n = rng.getrandbits(enc_bits) | 1
    g, x, y = (rng.getrandbits(enc_bits) for _ in range(3))
    recs = []
    orig_assign, orig_mulmod, orig_n2 = po.BigUintChip.assign_integer, po.BigUintChip.mul_mod, po.PaillierChip._n2

    def assign(self, ctx, v, bit_len):
        recs.append(("A", len(ctx.cells), bit_len // self.limb_bits, v))
        return orig_assign(self, ctx, v, bit_len)

    def mul_mod(self, ctx, a, b, nn, kind="mul"):
        first = len(ctx.cells)
        out = orig_mulmod(self, ctx, a, b, nn, kind)
        # q and rem are assigned through assign_integer inside mul_mod: drop those two records, the group covers them
        del recs[-2:]
        recs.append(("M", first, nn.num_limbs(), a.value, b.value, nn.value))
        return out

    def n2(self, ctx, pk):
        recs.append(("N", len(ctx.cells), pk.n.num_limbs(), pk.n.value))
        return orig_n2(self, ctx, pk)

    po.BigUintChip.assign_integer, po.BigUintChip.mul_mod, po.PaillierChip._n2 = assign, mul_mod, n2
    try:
        if add:
            res = po.paillier_add_native(n, x, y)
            ctx = po.paillier_enc_add_test(enc_bits, limb_bits, n, g, x, y, res, lookup_bits=lookup_bits)
        else:
            res = po.paillier_enc_native(n, g, x, y)
            ctx = po.paillier_enc_test(enc_bits, limb_bits, n, g, x, y, res, lookup_bits=lookup_bits)
    finally:
        po.BigUintChip.assign_integer, po.BigUintChip.mul_mod, po.PaillierChip._n2 = orig_assign, orig_mulmod, orig_n2
    cells = list(ctx.cells)
    if flip is not None:
        kind, which = flip
        target = [r for r in recs if r[0] == kind][which][1]
        cells[target] ^= 1
    with open(path, "w") as f:
        f.write(f"H {limb_bits} {lookup_bits}\n")
        for c in cells:
            f.write(f"C {c:x}\n")
        for r in recs:
            f.write(" ".join([r[0], str(r[1]), str(r[2])] + [f"{v:x}" for v in r[3:]]) + "\n")
    return len(cells), len(recs)


@pytest.mark.parametrize("enc_bits,limb_bits,add", [(128, 64, False), (264, 88, True), (264, 88, False)])
def test_cpp_constraint_checker_on_oracle_cells(tmp_path, enc_bits, limb_bits, add):
    """Host logic without a GPU: the C++ re-checker accepts the chip restatement's cell stream for the reference's two test flows
    and rejects it when one cell of an assign_integer, of the n^2 refresh or of a mul_mod group is flipped."""
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), CHECK_SRC, "-o", CHECK_EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    good = str(tmp_path / "flow.txt")
    n_cells, n_recs = _oracle_flow_file(good, enc_bits, limb_bits, 15, add=add)
    r = subprocess.run([CHECK_EXE, good], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "satisfied" in r.stdout, r.stdout + r.stderr
    for flip in (("A", 2), ("N", 0), ("M", -1), ("M", 0)):
        bad = str(tmp_path / f"bad_{flip[0]}{flip[1]}.txt")
        _oracle_flow_file(bad, enc_bits, limb_bits, 15, flip=flip, add=add)
        r = subprocess.run([CHECK_EXE, bad], capture_output=True, text=True, timeout=600)
        assert r.returncode == 3 and "violated" in r.stdout, (flip, r.stdout + r.stderr)
