import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from paillier_halo2_b200 import PaillierKey, workload
from paillier_halo2_b200.api import words_to_ints
for n_bits, count in ((2048, 4096), (1024, 8192)):
    kd = workload.load_key(n_bits)
    m_w, r_w = workload.units(n_bits, count)
    ms, rs = words_to_ints(m_w), words_to_ints(r_w)
    with PaillierKey(kd["n"], kd["g_rand"], n_bits, 64) as key:
        key.encrypt_witness_digest(ms[:64], rs[:64])
        t0 = time.perf_counter()
        cs, dig = key.encrypt_witness_digest(ms, rs)
        dt = time.perf_counter() - t0
        recs = sum(key.witness_records_for(m) for m in ms[:16]) / 16
        print(f"simple64 witness digest |n|={n_bits}: {count/dt:.1f} units/s, {recs:.0f} records/unit, {count*recs/dt/1e6:.2f} M mul_mod/s, {count*recs*2*key.words_out*8/dt/1e9:.2f} GB/s of witness")
