"""The built library really contains the Blackwell-native path: SASS of the block28u kernels carries tcgen05.mma (UTCIMMA), TMEM
loads / stores (LDTM / STTM), tcgen05.commit (UTCBAR) and TMEM allocation (UTCATOMSWS); the block28t kernels carry mma.sync (IMMA)
and none of those.  No GPU needed (cuobjdump disassembles the in-tree .so); skipped when cuobjdump is not on the box."""
import os
import re
import shutil
import subprocess

import pytest

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "paillier_halo2_b200", "libpaillier_b200.so")


def _cuobjdump():
    for c in (shutil.which("cuobjdump"), "/usr/local/cuda/bin/cuobjdump"):
        if c and os.path.exists(c):
            return c
    return None


@pytest.fixture(scope="module")
def sass(built_lib):
    exe = _cuobjdump()
    if not exe:
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, cur = {}, None
    for line in out.split("\n"):
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            per[cur] = {}
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
        if cur and m:
            per[cur][m.group(1)] = per[cur].get(m.group(1), 0) + 1
    return per


def _kernels(sass, name, shape, eng):
    # mangled: ...9k_encryptINS_3b283CfgILi8ELi19EEELi2EEEv...
    g, bl = shape
    pat = re.compile(rf"\d+{name}INS_3b283CfgILi{g}ELi{bl}EEELi{eng}EEEv")
    return [v for k, v in sass.items() if pat.search(k)]


@pytest.mark.parametrize("shape", [(8, 19), (16, 14), (16, 19)])
@pytest.mark.parametrize("name,eng", [("k_encrypt", 2), ("k_pow", 2), ("k_tally", 4), ("k_witness", 1), ("k_add_w", 1)])
def test_block28u_kernels_are_tcgen05(sass, shape, name, eng):
    ks = _kernels(sass, name, shape, eng)
    assert len(ks) == 1, (name, shape, eng, len(ks))
    ops = ks[0]
    assert ops.get("UTCIMMA", 0) >= 8 and ops.get("LDTM", 0) >= 2 and ops.get("STTM", 0) >= 1 and ops.get("UTCBAR", 0) >= 2
    assert ops.get("UTCATOMSWS", 0) >= 2          # tcgen05.alloc + dealloc
    assert ops.get("IMMA", 0) == 0                # no legacy mma.sync on this path
    assert ops.get("IMAD", 0) > 0                 # phase A: IMAD.WIDE products


@pytest.mark.parametrize("name,eng", [("k_encrypt", 1), ("k_witness", 0)])
def test_block28t_kernels_are_mma_sync(sass, name, eng):
    for shape in [(4, 19), (8, 19), (16, 14), (16, 19)]:
        ks = _kernels(sass, name, shape, eng)
        assert len(ks) == 1
        assert ks[0].get("IMMA", 0) > 0 and ks[0].get("UTCIMMA", 0) == 0 and ks[0].get("LDTM", 0) == 0


def test_no_tcgen05_variant_at_1024(sass):
    assert _kernels(sass, "k_encrypt", (4, 19), 2) == []
