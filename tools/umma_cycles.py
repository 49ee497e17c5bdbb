"""Cycles of phase A and of phases B + C per modular squaring, block28t (engine 3) against block28u (engine 4), one and two CTAs per SM.
modes: 0 token alternation (block28u), 1 no token, 16000 token + start-up stagger, -2 first CTA of an SM phase A only / second B + C only,
-3 phase A only, -4 phases B + C only.     python tools/umma_cycles.py [reps] [engines] -> JSON lines"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_halo2_b200 import PaillierKey, workload  # noqa: E402
from tools.umma_debug import image  # noqa: E402


def main():
    kd = workload.load_key(2048)
    key = PaillierKey(kd["n"], kd["g_rand"], 2048)
    lib = key._lib
    rng = np.random.default_rng(3)
    vals = [int.from_bytes(rng.bytes(525), "little") for _ in range(32)]
    v = image(vals, 8, 19)
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    engines = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3, 4, 5]
    for eng in engines:
        for ctas, mode in (((148, 1), (148, -3), (148, -4)) if eng == 5 else ((148, 1), (296, 1), (148, -3), (148, -4), (296, -3), (296, -4), (296, -2))):
            cyc = np.zeros((ctas, 3), dtype=np.int64)
            rc = lib.pb200_debug_mulmod_cycles(key.handle, eng, v.ctypes.data, ctas, reps, mode, cyc.ctypes.data)
            nz = lambda col: float(col[col > 0].mean()) / reps if (col > 0).any() else 0.0  # noqa: E731
            a, b, t = nz(cyc[:, 0]), nz(cyc[:, 1]), nz(cyc[:, 2])
            print(json.dumps({"engine": eng, "ctas": ctas, "mode": mode, "rc": rc, "phaseA_clk": round(a), "phasesBC_clk": round(b),
                              "loop_clk_per_mulmod": round(t), "sm_clk_per_lane_mulmod": round(t / (64 if eng == 5 else 32) / (ctas / 148), 1)}))


if __name__ == "__main__":
    main()
