"""Generate tests/golden/cells.json: advice-cell streams (K4) from the chip restatement in oracle/paillier_oracle.py.

The reference holds no fixtures for its cells (layout "parity unpinned"); these pin OUR restatement so that
the oracle and the GPU kernels cannot drift together unnoticed.  A stream is hashed as SHA-256 over its cells, each written
as 32 little-endian bytes.  Re-run: `python tools/gen_golden_cells.py`.

`python tools/gen_golden_cells.py --dump IDX` writes the cells of flow IDX of the committed fixture, one 32-byte little-endian hex
line per cell, to stdout: the file to diff against `target/advice_cells_flowIDX.hex` of rust/paillier-b200/tests/mockprover_cells.rs
(the pin-on-first-toolchain recipe)."""
import hashlib, json, os, random, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.paillier_oracle import Assigned, BigUintChip, Context, decompose, paillier_enc_native, paillier_enc_test
from paillier_halo2_b200 import workload


def cell_hash(cells):
    h = hashlib.sha256()
    for c in cells:
        h.update(int(c).to_bytes(32, "little"))
    return h.hexdigest()


if len(sys.argv) == 3 and sys.argv[1] == "--dump":
    f = json.load(open(os.path.join(ROOT, "tests", "golden", "cells.json")))["flows"][int(sys.argv[2])]
    hx = lambda k: int(f[k], 16)
    ctx = paillier_enc_test(f["enc_bits"], f["limb_bits"], hx("n"), hx("g"), hx("m"), hx("r"), hx("c"), lookup_bits=f["lookup_bits"])
    assert cell_hash(ctx.cells) == f["sha256"]
    for c in ctx.cells:
        print(int(c).to_bytes(32, "little").hex())
    sys.exit(0)

rng = random.Random(0xCE115)
out = {"flows": [], "groups": []}
for enc_bits, limb_bits, lookup_bits in ((128, 64, 15), (264, 88, 15), (128, 64, 13)):
    n = rng.getrandbits(enc_bits) | (1 << (enc_bits - 1)) | 1
    g = rng.getrandbits(enc_bits)
    for m, r in ((rng.getrandbits(enc_bits), rng.getrandbits(enc_bits)), (0, rng.getrandbits(enc_bits)), (1, 1)):
        c = paillier_enc_native(n, g, m, r)
        ctx = paillier_enc_test(enc_bits, limb_bits, n, g, m, r, c, lookup_bits=lookup_bits)
        out["flows"].append({"enc_bits": enc_bits, "limb_bits": limb_bits, "lookup_bits": lookup_bits, "n": hex(n), "g": hex(g),
                             "m": hex(m), "r": hex(r), "c": hex(c), "n_cells": len(ctx.cells), "sha256": cell_hash(ctx.cells),
                             "first": [hex(v) for v in ctx.cells[:6]], "last": [hex(v) for v in ctx.cells[-6:]]})
for n_bits in (1024, 2048):
    n = workload.load_key(n_bits)["n"]
    n2 = n * n
    L = 2 * n_bits // 64
    for a, b in ((rng.randrange(n2), rng.randrange(n2)), (n2 - 1, n2 - 1)):
        q, rem = divmod(a * b, n2)
        ctx = Context()
        BigUintChip(64, 15).mul_mod(ctx, Assigned(decompose(a, L, 64), a, 64), Assigned(decompose(b, L, 64), b, 64),
                                    Assigned(decompose(n2, L, 64), n2, 64))
        out["groups"].append({"n_bits": n_bits, "lookup_bits": 15, "a": hex(a), "b": hex(b), "q": hex(q), "rem": hex(rem),
                              "n_cells": len(ctx.cells), "sha256": cell_hash(ctx.cells)})
path = os.path.join(ROOT, "tests", "golden", "cells.json")
json.dump(out, open(path, "w"), indent=1)
print(path, len(out["flows"]), "flows,", len(out["groups"]), "groups")
