"""Digit-for-digit comparison of the block28 engines on one CTA's modular multiplication (GPU, through the C ABI).

block28t (mma.sync) and block28u (tcgen05) are specified to leave IDENTICAL lazy digits.  This tool feeds random lazy values through
pb200_debug_mulmod on both and reports, stage by stage, where they differ: the 2L-digit product of phase A (TMEM stash), the packed
q-hat rows of phase B, the value after phase C; then one small encrypt batch on every engine against each other.
    python tools/umma_debug.py [n_bits] -> one JSON line, exit code 1 on any mismatch
"""
import ctypes as C
import json
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_halo2_b200 import PaillierKey, workload  # noqa: E402
from paillier_halo2_b200 import _lib  # noqa: E402

W = 28


def image(values, G, BL):
    """lazy values (python ints, signed) of 32 lanes -> shared-memory image [block][chunk][lane][4] int32 of strict centred digits"""
    CH = (BL + 3) // 4
    img = np.zeros((G, CH, 32, 4), dtype=np.int32)
    for lane, v in enumerate(values):
        carry = 0
        for p in range(G * BL):
            t = ((v >> (W * p)) & ((1 << W) - 1)) + carry
            d = ((t + (1 << (W - 1))) & ((1 << W) - 1)) - (1 << (W - 1))
            carry = (t - d) >> W
            img[p // BL, (p % BL) // 4, lane, (p % BL) % 4] = d
    return img


def value_of(img, lane, G, BL, blocks=None):
    v = 0
    nb = blocks or G
    for p in range(nb * BL):
        v += int(img[p // BL, (p % BL) // 4, lane, (p % BL) % 4]) << (W * p)
    return v


def main():
    n_bits = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    kd = workload.load_key(n_bits)
    key = PaillierKey(kd["n"], kd["g_rand"], n_bits)
    lib = key._lib
    g, bl = C.c_int(), C.c_int()
    assert lib.pb200_key_shape(key.handle, C.byref(g), C.byref(bl)) == 0
    G, BL = g.value, bl.value
    L, CH = G * BL, (BL + 3) // 4
    rng = np.random.default_rng(7)
    beta = 14 * (2 * L - 1)

    def rnd_vals():
        out = []
        for lane in range(32):
            v = int.from_bytes(rng.bytes((beta - 2 + 7) // 8), "little") >> ((8 - (beta - 2) % 8) % 8)
            if lane % 3 == 1:
                v = -v
            if lane == 5:
                v = 0
            if lane == 6:
                v = 1
            if lane == 7:
                v = (1 << (beta - 2)) - 1
            out.append(v)
        return out

    res = {"n_bits": n_bits, "G": G, "BL": BL, "engine_default": key.engine}
    bad = False
    for mode in ("sqr", "mul"):
        v = image(rnd_vals(), G, BL)
        y = image(rnd_vals(), G, BL) if mode == "mul" else None
        outs = {}
        for eng in (3, 4, 5):
            t = np.zeros((2 * G, CH, 32, 4), dtype=np.int32)
            rc = lib.pb200_debug_mulmod(key.handle, eng, v.ctypes.data, y.ctypes.data if y is not None else None, 1, None, t.ctypes.data, None)
            if rc == _lib.PB200_ERR_UNSUPPORTED:       # this key size has no such variant
                continue
            if rc:
                res[f"{mode}_phaseA_rc_eng{eng}"] = rc
                bad = True
                continue
            vo = np.zeros_like(v)
            rows = np.zeros((32, L), dtype=np.uint32)
            rc = lib.pb200_debug_mulmod(key.handle, eng, v.ctypes.data, y.ctypes.data if y is not None else None, 1, vo.ctypes.data, None, rows.ctypes.data)
            v5 = np.zeros_like(v)
            rc5 = lib.pb200_debug_mulmod(key.handle, eng, v.ctypes.data, y.ctypes.data if y is not None else None, 5, v5.ctypes.data, None, None)
            if rc or rc5:
                res[f"{mode}_rc_eng{eng}"] = [rc, rc5]
                bad = True
                continue
            outs[eng] = (t, vo, rows, v5)
        for other in (4, 5):
          if 3 in outs and other in outs:
            for name, idx in (("phaseA_T", 0), ("v_out", 1), ("qhat_rows", 2), ("v_out_5reps", 3)):
                name = f"eng{other}_{name}"
                a, b = outs[3][idx], outs[other][idx]
                neq = a != b
                n = int(neq.sum())
                res[f"{mode}_{name}_mismatch"] = n
                if n:
                    bad = True
                    where = np.argwhere(neq)
                    res[f"{mode}_{name}_first"] = [[int(x) for x in w] for w in where[:6]]
                    res[f"{mode}_{name}_a"] = [int(a[tuple(w)]) for w in where[:6]]
                    res[f"{mode}_{name}_b"] = [int(b[tuple(w)]) for w in where[:6]]
                    if name.endswith("qhat_rows"):
                        res[f"{mode}_qhat_bad_digits"] = sorted(set(int(w[1]) for w in where))[:40]
                        res[f"{mode}_qhat_bad_lanes"] = sorted(set(int(w[0]) for w in where))[:40]
        if 3 in outs:
            # the product itself against Python: T == v * y for lane 0..3 (engine 3 is the reference of the comparison)
            t3 = outs[3][0]
            ok = True
            vv = [value_of(v, ln, G, BL) for ln in range(4)]
            yy = [value_of(y, ln, G, BL) for ln in range(4)] if y is not None else vv
            for ln in range(4):
                if value_of(t3, ln, 2 * G, BL) != vv[ln] * yy[ln]:
                    ok = False
            res[f"{mode}_phaseA_vs_python"] = ok
            bad |= not ok
    # small encrypt batch on every fast engine
    m, r = workload.units(n_bits, 101)
    cs = {}
    for eng in (2, 3, 4, 5):
        try:
            key.set_engine(eng)
            cs[eng] = key.encrypt_words(m, r)
        except Exception as e:  # noqa: BLE001
            if getattr(e, "status", None) == _lib.PB200_ERR_UNSUPPORTED:
                continue
            res[f"encrypt_eng{eng}_error"] = str(e)[:200]
            bad = True
    for other in (4, 5):
        if 3 in cs and other in cs:
            res[f"encrypt_eng{other}_vs_eng3_mismatch_units"] = int((cs[3] != cs[other]).any(axis=1).sum())
            bad |= res[f"encrypt_eng{other}_vs_eng3_mismatch_units"] != 0
    if 2 in cs and 3 in cs:
        res["encrypt_eng3_vs_eng2_mismatch_units"] = int((cs[2] != cs[3]).any(axis=1).sum())
    res["ok"] = not bad
    print(json.dumps(res))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
