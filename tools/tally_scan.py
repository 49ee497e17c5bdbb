"""Time of the single-launch tally (pb200_tally_dev) against the number of ciphertexts on one GPU: separates the fixed tail
(tree over the CTAs, lane fold, finalize) from the streaming part.   python tools/tally_scan.py [engine] -> JSON lines"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paillier_halo2_b200 import PaillierKey, workload  # noqa: E402


def main():
    eng = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    kd = workload.load_key(2048)
    key = PaillierKey(kd["n"], kd["g_rand"], 2048)
    if eng:
        key.set_engine(eng)
    cs = workload.ciphertexts(2048, 1 << 20, kd["n"])
    d_c = torch.from_numpy(cs.view(np.int64)).cuda()
    out = torch.empty(key.words_out, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.ExternalStream(key.stream)
    for count in (32, 1024, 9472, 1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
        ts = []
        for it in range(6):
            flush.fill_(it)
            torch.cuda.synchronize()
            with torch.cuda.stream(st):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                key.tally_dev(d_c.data_ptr(), count, out.data_ptr())
                e1.record(st)
            key.sync()
            ts.append(e0.elapsed_time(e1))
        print(json.dumps({"engine": key.engine, "count": count, "ms_min": round(min(ts[1:]), 4), "ms_med": round(float(np.median(ts[1:])), 4)}))


if __name__ == "__main__":
    main()
