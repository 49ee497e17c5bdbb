"""Throughput of the K4 cell-expansion kernel (k_cells_mulmod): mul_mod groups/s and GB/s of cells written to HBM."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from paillier_halo2_b200 import PaillierKey, workload

def run(n_bits, count, lookup, mont, reps=3):
    kd = workload.load_key(n_bits)
    dev = torch.device("cuda:0")
    with PaillierKey(kd["n"], kd["g_std"], n_bits, 64) as key:
        wo = key.words_out
        c_w = workload.ciphertexts(n_bits, 2 * count, kd["n"])
        d_a = torch.from_numpy(c_w[:count].view(np.int64)).to(dev); d_b = torch.from_numpy(c_w[count:].view(np.int64)).to(dev)
        d_rem = torch.empty_like(d_a); d_q = torch.empty_like(d_a)
        key.add_dev(d_a.data_ptr(), d_b.data_ptr(), wo, count, d_rem.data_ptr(), d_q.data_ptr()); key.sync()
        for eng in (0, 1):
            key.set_engine(eng)
            key.add_dev(d_a.data_ptr(), d_b.data_ptr(), wo, min(count, 4096), d_rem.data_ptr(), d_q.data_ptr()); key.sync()
            t0 = time.perf_counter()
            key.add_dev(d_a.data_ptr(), d_b.data_ptr(), wo, count, d_rem.data_ptr(), d_q.data_ptr()); key.sync()
            dt = time.perf_counter() - t0
            print(json.dumps({"n_bits": n_bits, "add_with_quotient": key.witness_engine, "pairs_per_s": count / dt,
                              "hbm_GBps": count * 4 * wo * 8 / dt / 1e9}), flush=True)
        key.set_engine(0)
        per = key.cells_layout(lookup)["cells_per_mulmod"]
        d_cells = torch.empty((count, per, 4), dtype=torch.int64, device=dev)
        stream = torch.cuda.ExternalStream(key.stream, device=dev)
        key.mulmod_cells_dev(d_a.data_ptr(), d_b.data_ptr(), d_q.data_ptr(), d_rem.data_ptr(), count, lookup, mont, d_cells.data_ptr()); key.sync()
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            key.mulmod_cells_dev(d_a.data_ptr(), d_b.data_ptr(), d_q.data_ptr(), d_rem.data_ptr(), count, lookup, mont, d_cells.data_ptr())
            e1.record(stream); key.sync(); ms.append(e0.elapsed_time(e1))
        t = min(ms) * 1e-3
        out = {"n_bits": n_bits, "groups": count, "lookup_bits": lookup, "montgomery": bool(mont), "cells_per_group": per,
               "groups_per_s": count / t, "cells_GBps": count * per * 32 / t / 1e9, "ms": min(ms)}
        print(json.dumps(out), flush=True)
        return out

if __name__ == "__main__":
    if len(sys.argv) >= 5:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), reps=1); sys.exit(0)
    res = [run(2048, 65536, 15, 0), run(2048, 65536, 15, 1), run(2048, 65536, 0, 0), run(1024, 131072, 15, 0), run(3072, 32768, 15, 0), run(4096, 16384, 15, 0)]
    json.dump(res, open("/root/repo/gpurun_out/cells_bench.json", "w"), indent=1)
