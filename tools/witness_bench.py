"""Throughput of the witness path (pb200_encrypt_witness_digest_dev, inputs resident in HBM): units/s and mul_mod/s."""
import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
from paillier_halo2_b200 import PaillierKey, workload

def run(n_bits, count, engine):
    kd = workload.load_key(n_bits)
    m_w, r_w = workload.units(n_bits, count)
    dev = torch.device("cuda:0")
    d_m = torch.from_numpy(m_w.view(np.int64)).to(dev); d_r = torch.from_numpy(r_w.view(np.int64)).to(dev)
    with PaillierKey(kd["n"], kd["g_rand"], n_bits, 64) as key:
        key.set_engine(engine)
        wo = key.words_out
        d_c = torch.empty((count, wo), dtype=torch.int64, device=dev)
        d_d = torch.empty(count, dtype=torch.int64, device=dev)
        torch.cuda.synchronize()
        key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), min(count, 64), d_c.data_ptr(), d_d.data_ptr()); key.sync()
        t0 = time.perf_counter()
        key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), count, d_c.data_ptr(), d_d.data_ptr()); key.sync()
        dt = time.perf_counter() - t0
        recs = float(np.mean([key.witness_records_for(int.from_bytes(m_w[i].tobytes(), "little")) for i in range(16)]))
        out = {"n_bits": n_bits, "count": count, "engine": key.witness_engine, "arithmetic": key.engine, "units_per_s": count / dt, "records_per_unit": recs,
               "mul_mod_per_s": count * recs / dt, "witness_GBps": count * recs * 2 * wo * 8 / dt / 1e9, "seconds": dt}
        print(json.dumps(out), flush=True)
        return out

if __name__ == "__main__":
    if len(sys.argv) >= 3:      # one configuration: n_bits count [engine] (ncu target; engine 4 = tcgen05 phases, 3 = mma.sync)
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 0)
        sys.exit(0)
    res = []
    for n_bits, count in ((2048, 65536), (1024, 65536), (3072, 16384), (4096, 8192)):
        res.append(run(n_bits, count, 0))
    res.append(run(2048, 2048, 1))
    json.dump(res, open("/root/repo/gpurun_out/witness_bench.json", "w"), indent=1)
