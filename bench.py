#!/usr/bin/env python
"""bench.py — Paillier enc/s at |n|=2048 on N B200s (BASELINE.json metric), one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W            # our arm   (N>1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU path on the host cores

A step is one pass of the hot path over one batch: encrypt 2^16 independent (m, r) pairs under one
2048-bit key (BASELINE.json configs[1]; SURVEY.md §8d inputs: seeded Paillier key, random g in
[2, 2^|n|), full-width Philox m, r).  `value` times the device-resident call (inputs already in HBM) with
CUDA events on the stream the kernel is launched on; `e2e` times the host-buffer C-ABI call
(pb200_encrypt_batch: pinned host -> device, kernel, device -> host) by wall clock around the blocking call.
Multi-GPU: units are independent, so each rank encrypts its own 2^16 units (weak scaling, no data-path
collective); the time is the max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  NCCL prints its version banner to file descriptor 1 when the first communicator is
# created, so everything except the result line is sent to stderr: fd 1 is duplicated for the result, then pointed at fd 2.
_RESULT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict) -> None:
    _RESULT.write(json.dumps(line) + "\n")
    _RESULT.flush()


N_BITS = 2048
UNITS = 1 << 16
# dram__bytes_read.sum + dram__bytes_write.sum of one k_encrypt launch (ncu --set full, profiles/), keyed by (|n|, units)
TRAFFIC_BYTES = {(2048, 1 << 16): 6121053000 + 806334464}   # profiles/ncu_k_encrypt_r01_fused_summary.txt
METRIC = "paillier_enc_per_s_n2048"
UNIT = "enc/s"


def mac_counts(n_bits: int):
    """Algorithmic 32x32->64 MACs per modular multiplication / squaring over n^2:
    W_mul = 2*L32^2 + L32, W_sqr = (L32^2 + L32)/2 + L32^2 + L32, L32 = 2|n|/32."""
    l32 = 2 * n_bits // 32
    return 2 * l32 * l32 + l32, (l32 * l32 + l32) // 2 + l32 * l32 + l32


def imad_peak():
    """P_imad: measured rate of 32x32->64 multiplies (SASS IMAD.WIDE) on this pool's B200, from
    profiles/imad_peak2_r01.json (csrc/microbench/imad_peak2.cu, variant imadw_only: 31.9 per clk per SM,
    half the 32-bit IMAD rate).  MEASURED_PEAKS.json carries no integer-pipe peak."""
    path = os.path.join(ROOT, "profiles", "imad_peak2_r01.json")
    try:
        d = json.load(open(path))
        return max(r["op_per_s"] for r in d["results"] if r["variant"] == "imadw_only"), "profiles/imad_peak2_r01.json (IMAD.WIDE, measured)"
    except Exception:
        return 148 * 32 * 1.965e9, "nominal 148 SM x 32 IMAD.WIDE/clk x 1.965 GHz"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(threads: int, units: int, key: dict):
    """The CPU restatement (oracle/paillier_cpu.cpp: OpenSSL BIGNUM port of src/paillier.rs:87-92) on the
    first `units` units of the same workload, `threads` worker threads.  Returns (enc/s, seconds)."""
    from oracle import cpu_ref
    from paillier_halo2_b200 import workload

    m_w, r_w = workload.units(N_BITS, units)
    t0 = time.perf_counter()
    cpu_ref.enc_batch(key["n"], key["g_rand"], N_BITS // 64, m_w, r_w, threads=threads)
    dt = time.perf_counter() - t0
    return units / dt, dt


def run_reference(args):
    """Reference arm: the reference's own CPU algorithm (OpenSSL port; the Rust original cannot be built
    here) with every host thread, on bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_ref
    from paillier_halo2_b200 import workload

    key = workload.load_key(N_BITS)
    threads = cpu_ref.hardware_threads()
    sample = max(64, min(UNITS, 96 * threads))          # ~3 s of CPU work per step at ~35 enc/s/thread
    m_w, r_w = workload.units(N_BITS, sample)
    for _ in range(args.warmup):
        cpu_ref.enc_batch(key["n"], key["g_rand"], N_BITS // 64, m_w[: max(threads, 8)], r_w[: max(threads, 8)], threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_ref.enc_batch(key["n"], key["g_rand"], N_BITS // 64, m_w, r_w, threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (OpenSSL BIGNUM)", "data": "synthetic",
        "config": {"workload": f"batched encrypt |n|={N_BITS}, bounded sample of {sample} units/step of the 2^16-unit batch, random g",
                   "key": "seeded p*q (SURVEY.md 8d)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} units x {args.steps} steps, OpenSSL BN_mod_exp x2 + BN_mod_mul per unit"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def hbm_peak():
    """Measured HBM copy bandwidth of this pool's B200 (driver-written MEASURED_PEAKS.json), else the profiling guide's fallback."""
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def cells_leg(key, kd, n_bits, groups=1 << 16, lookup_bits=15, reps=4):
    """K4: advice cells of `groups` mul_mod groups, device-resident, CUDA events on the key's stream.
    Algorithmic bytes per group = cells_per_mulmod x 32 B written + 4 x words_out x 8 B read."""
    import numpy as np
    import torch
    from paillier_halo2_b200 import workload
    dev = torch.device("cuda", key.device)
    wo = key.words_out
    c_w = workload.ciphertexts(n_bits, 2 * groups, kd["n"])
    d_a = torch.from_numpy(c_w[:groups].view(np.int64)).to(dev)
    d_b = torch.from_numpy(c_w[groups:].view(np.int64)).to(dev)
    d_rem, d_q = torch.empty_like(d_a), torch.empty_like(d_a)
    key.add_dev(d_a.data_ptr(), d_b.data_ptr(), wo, groups, d_rem.data_ptr(), d_q.data_ptr())
    key.sync()
    per = key.cells_layout(lookup_bits)["cells_per_mulmod"]
    d_cells = torch.empty((groups, per, 4), dtype=torch.int64, device=dev)       # 5.3 GB at |n| = 2048: larger than L2
    stream = torch.cuda.ExternalStream(key.stream, device=dev)
    peak, src = hbm_peak()
    out = {"groups": groups, "lookup_bits": lookup_bits, "cells_per_group": per}
    for mont in (0, 1):
        for _ in range(3):
            key.mulmod_cells_dev(d_a.data_ptr(), d_b.data_ptr(), d_q.data_ptr(), d_rem.data_ptr(), groups, lookup_bits, bool(mont), d_cells.data_ptr())
        key.sync()
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            key.mulmod_cells_dev(d_a.data_ptr(), d_b.data_ptr(), d_q.data_ptr(), d_rem.data_ptr(), groups, lookup_bits, bool(mont), d_cells.data_ptr())
            e1.record(stream)
            key.sync()
            ms.append(e0.elapsed_time(e1))
        t = sum(ms) / len(ms) * 1e-3
        gbs = groups * (per * 32 + 4 * wo * 8) / t / 1e9
        out["montgomery" if mont else "canonical"] = {"groups_per_s": groups / t, "ms_per_launch": t * 1e3,
                                                      "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                                                   "peak_source": src}}
    # parity of one group against the chip restatement (checker only)
    from oracle.paillier_oracle import Assigned, BigUintChip, Context, decompose
    from paillier_halo2_b200.api import _cells_to_ints, words_to_ints
    L = 2 * n_bits // 64
    a0, b0 = words_to_ints(c_w[:1])[0], words_to_ints(c_w[groups:groups + 1])[0]
    n2 = kd["n"] ** 2
    key.mulmod_cells_dev(d_a.data_ptr(), d_b.data_ptr(), d_q.data_ptr(), d_rem.data_ptr(), 1, lookup_bits, False, d_cells.data_ptr())
    key.sync()
    got = _cells_to_ints(d_cells[0].cpu().numpy().view(np.uint64))
    ctx = Context()
    BigUintChip(64, lookup_bits).mul_mod(ctx, Assigned(decompose(a0, L, 64), a0, 64), Assigned(decompose(b0, L, 64), b0, 64),
                                         Assigned(decompose(n2, L, 64), n2, 64))
    out["parity"] = bool(got == ctx.cells)
    del d_cells
    return out


def run_tally(args):
    """BASELINE.json configs[2]: product of 2^20 ciphertexts mod n^2, sharded over the GPUs with one all-gather of
    the per-GPU partials (NCCL) and a final combine.  Strong scaling (total work fixed)."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from paillier_halo2_b200 import PaillierKey, _lib, workload
    from paillier_halo2_b200.shard import shard_range, tally_sharded_gpu

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    kd = workload.load_key(N_BITS)
    total = 1 << 20
    key = PaillierKey(kd["n"], kd["g_std"], N_BITS, 64, device=local_rank)
    lo, hi = shard_range(total, rank, world)
    c_all = workload.ciphertexts(N_BITS, total, kd["n"]) if total <= (1 << 20) else None
    d_c = torch.from_numpy(c_all[lo:hi].view(np.int64)).cuda()
    stream = torch.cuda.ExternalStream(key.stream, device=torch.device("cuda", local_rank))
    partial = torch.empty(key.words_out, dtype=torch.int64, device="cuda")
    for _ in range(args.warmup):
        out = tally_sharded_gpu(key, d_c, hi - lo, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = lib.pb200_kernel_launches()
    kern_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); key.tally_dev(d_c.data_ptr(), hi - lo, partial.data_ptr()); e1.record(stream)
        key.sync(); kern_ms += e0.elapsed_time(e1)
        out = tally_sharded_gpu(key, d_c, hi - lo, world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = lib.pb200_kernel_launches() - launches0
    t = torch.tensor([dt, kern_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt, kern_ms = float(t[0]), float(t[1])
    if rank == 0:
        from oracle import cpu_ref
        want = cpu_ref.tally(kd["n"], N_BITS // 64, c_all, threads=cpu_ref.hardware_threads())
        ok = bool((out.cpu().numpy().view(np.uint64) == want).all())
        w_mul, _ = mac_counts(N_BITS)
        peak, src = imad_peak()
        shard = hi - lo
        ksec = kern_ms * 1e-3 / args.steps
        line = {"metric": f"paillier_tally_ciphertexts_per_s_n{N_BITS}", "value": total * 2 * args.steps / dt, "unit": "ciphertexts/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / (2 * args.steps),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64 accumulators over signed 28-bit digits",
                "data": "synthetic", "config": {"workload": f"product of 2^20 ciphertexts mod n^2, |n|={N_BITS}, sharded over {world} GPU(s), "
                                                            "NCCL all-gather of partials + combine (BASELINE.json configs[2])", "engine": key.engine},
                "gpu_launches": int(launches), "parity_vs_cpu_fold": ok,
                "roofline": {"bound": "imad", "achieved": shard * w_mul / ksec / 1e12, "peak": peak / 1e12, "unit": "TMAC/s",
                             "frac": shard * w_mul / ksec / peak, "peak_source": src,
                             "hbm_gbs_achieved": shard * key.words_out * 8 / ksec / 1e9,
                             "note": "per-GPU shard fold kernel (k_tally x2 launches); one modmul per 512 B read: compute-bound, HBM GB/s reported because north_star asks for it"}}
        emit(line)
    key.close()
    if world > 1:
        dist.destroy_process_group()


def run_witness(args):
    """BASELINE.json configs[3]: batched encrypt + the (q, rem) limb witness of every mul_mod PaillierChip::encrypt issues
    (|n| = 3072 and 2^18 units over 8 GPUs in the config; --n-bits / --units select others).  Weak scaling: each rank runs
    `units` units of its own slice; no collective on the data path.  Value = units/s, whole job."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from paillier_halo2_b200 import PaillierKey, _lib, workload
    from paillier_halo2_b200.api import witness_digest, words_to_ints

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    kd = workload.load_key(N_BITS)
    units = args.units
    key = PaillierKey(kd["n"], kd["g_rand"], N_BITS, 64, device=local_rank)
    m_w, r_w = workload.units(N_BITS, units, seed_offset=1000 * rank)
    d_m = torch.from_numpy(m_w.view(np.int64)).cuda(); d_r = torch.from_numpy(r_w.view(np.int64)).cuda()
    d_c = torch.empty((units, key.words_out), dtype=torch.int64, device="cuda")
    d_dig = torch.empty(units, dtype=torch.int64, device="cuda")
    stream = torch.cuda.ExternalStream(key.stream, device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), min(units, 9472), d_c.data_ptr(), d_dig.data_ptr())
    for _ in range(max(args.warmup - 1, 0)):
        key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), min(units, 9472), d_c.data_ptr(), d_dig.data_ptr())
    sampler = ClockSampler(local_rank)
    launches0 = lib.pb200_kernel_launches()
    barrier()
    sampler.start()
    evs = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), units, d_c.data_ptr(), d_dig.data_ptr())
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    launches = lib.pb200_kernel_launches() - launches0
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    if rank == 0:
        from oracle.paillier_oracle import encrypt_steps
        w_mul, w_sqr = mac_counts(N_BITS)
        pop_n = bin(kd["n"]).count("1")
        pop_m = float(np.mean([bin(v).count("1") for v in words_to_ints(m_w[:256])]))
        n_sqr, n_mul = kd["n"].bit_length(), pop_n + pop_m + 1
        a_wit = n_sqr * w_sqr + n_mul * w_mul
        dig = d_dig.cpu().numpy().view(np.uint64)
        cs = d_c.cpu().numpy().view(np.uint64)
        ok = True
        for i in sorted({0, units - 1} | {(units * t) // 16 for t in range(16)}):       # 17 units spread over the batch
            mi, ri = words_to_ints(m_w[i:i + 1])[0], words_to_ints(r_w[i:i + 1])[0]
            c, steps = encrypt_steps(kd["n"], kd["g_rand"], mi, ri)
            gs = mi.bit_length() + bin(mi).count("1")
            mine = [(x.q, x.rem) for x in steps[:gs] if x.kind == "mul"] + [(x.q, x.rem) for x in steps[gs:]]
            ok = ok and int(dig[i]) == witness_digest(mine, key.words_out) and words_to_ints(cs[i:i + 1])[0] == c
        peak, src = imad_peak()
        ksec = dev_ms * 1e-3 / args.steps
        line = {"metric": f"paillier_witness_units_per_s_n{N_BITS}", "value": world * units * args.steps / (dev_ms * 1e-3), "unit": "units/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int64 columns over signed 28-bit digits + s8 IMMA, exact (q, rem) tail",
                "data": "synthetic",
                "config": {"workload": f"batched encrypt + (q, rem) limb witness of every mul_mod, |n|={N_BITS}, {units} units per GPU per step "
                                       "(BASELINE.json configs[3]), witness digested on the device", "engine": key.witness_engine,
                           "chain": {"mod_sqr": n_sqr, "mod_mul": n_mul, "mac_per_unit": a_wit}},
                "gpu_launches": int(launches), "clocks": clocks, "parity": ok,
                "mul_mod_per_s": world * units * args.steps * (n_sqr + n_mul) / (dev_ms * 1e-3),
                "roofline": {"bound": "imad", "achieved": units * a_wit / ksec / 1e12, "peak": peak / 1e12, "unit": "TMAC/s",
                             "frac": units * a_wit / ksec / peak, "peak_source": src, "traffic": None}}
        emit(line)
    key.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--units", type=int, default=UNITS, help="units per GPU per step (default: the BASELINE config, 2^16)")
    ap.add_argument("--g", default="rand", choices=["rand", "std"], help="rand: random g (headline); std: g = n+1")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-witness", action="store_true", help="skip the witness-mode leg")
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--n-bits", type=int, default=2048, help="key size |n| (default 2048, the BASELINE metric; others are the sweep)")
    ap.add_argument("--workload", default="encrypt", choices=["encrypt", "tally", "witness"],
                    help="encrypt: the BASELINE metric (default); tally: product of 2^20 ciphertexts sharded over the GPUs (configs[2]); "
                         "witness: encrypt + (q, rem) witness of every mul_mod (configs[3]: --n-bits 3072 --units 32768 on 8 GPUs)")
    args = ap.parse_args()
    global N_BITS, METRIC
    N_BITS = args.n_bits
    METRIC = f"paillier_enc_per_s_n{N_BITS}"
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "tally":
        return run_tally(args)
    if args.workload == "witness":
        return run_witness(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    from paillier_halo2_b200 import PaillierKey, _lib, workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the Paillier hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    kd = workload.load_key(N_BITS)
    g = kd["g_rand"] if args.g == "rand" else kd["g_std"]
    units = args.units
    key = PaillierKey(kd["n"], g, N_BITS, 64, device=local_rank)
    if args.engine:
        key.set_engine(args.engine)
    n_sqr, n_mul = key.chain_counts()
    w_mul, w_sqr = mac_counts(N_BITS)
    a_enc = n_sqr * w_sqr + n_mul * w_mul

    # this rank's units: rank r takes the r-th 2^16-unit slice of the Philox stream
    m_w, r_w = workload.units(N_BITS, units, seed_offset=1000 * rank)
    m_pin = torch.from_numpy(m_w.view(np.int64)).pin_memory()
    r_pin = torch.from_numpy(r_w.view(np.int64)).pin_memory()
    c_pin = torch.empty((units, key.words_out), dtype=torch.int64).pin_memory()
    d_m = m_pin.cuda(non_blocking=False)
    d_r = r_pin.cuda(non_blocking=False)
    d_c = torch.empty((units, key.words_out), dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    stream = torch.cuda.ExternalStream(key.stream, device=torch.device("cuda", local_rank))

    def step_dev():
        key.encrypt_dev(d_m.data_ptr(), d_r.data_ptr(), units, d_c.data_ptr())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step_dev()
    key.sync()

    # ---- timed: K steps, device-resident inputs, CUDA events on the launching stream, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    launches0 = lib.pb200_kernel_launches()
    barrier()
    sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_dev()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    clocks = sampler.stop()
    launches = lib.pb200_kernel_launches() - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    value = world * units * args.steps / (dev_ms * 1e-3)

    # ---- e2e: the host-buffer C-ABI call, pinned host memory, H2D + kernel + D2H inside the timed region
    import ctypes as C
    hp = [C.cast(t_.data_ptr(), _lib.u64p) for t_ in (m_pin, r_pin, c_pin)]

    def step_host():
        _lib.check(lib.pb200_encrypt_batch(key.handle, hp[0], hp[1], units, hp[2]), "pb200_encrypt_batch")

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * units * args.steps / e2e_s

    # ---- witness mode (BASELINE.json configs[1]: "witnesses checked bit-exact"): the reference's own LSB-first chain with the
    # exact (q, rem) of every mul_mod, digested on the device (pb200_encrypt_witness_digest_dev), same units, inputs in HBM
    witness = None
    if not args.no_witness:
        d_dig = torch.empty(units, dtype=torch.int64, device="cuda")
        d_cw = torch.empty((units, key.words_out), dtype=torch.int64, device="cuda")
        key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), units, d_cw.data_ptr(), d_dig.data_ptr())     # warm-up at full size
        barrier()
        wl0 = lib.pb200_kernel_launches()
        w_sampler = ClockSampler(local_rank)
        w_sampler.start()
        w_evs = []
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            key.encrypt_witness_digest_dev(d_m.data_ptr(), d_r.data_ptr(), units, d_cw.data_ptr(), d_dig.data_ptr())
            e1.record(stream)
            w_evs.append((e0, e1))
        barrier()
        w_clocks = w_sampler.stop()
        w_ms = sum(a.elapsed_time(b) for a, b in w_evs) / len(w_evs)
        t = torch.tensor([w_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w_ms = float(t.item())
        if rank == 0:
            from oracle.paillier_oracle import encrypt_steps
            from paillier_halo2_b200.api import witness_digest, words_to_ints
            pop_n = bin(kd["n"]).count("1")
            pop_m = float(np.mean([bin(v).count("1") for v in words_to_ints(m_w[:256])]))
            w_sqr_n, w_mul_n = kd["n"].bit_length(), pop_n + pop_m + 1
            a_wit = w_sqr_n * w_sqr + w_mul_n * w_mul
            ok = bool((d_cw.cpu().numpy().view(np.uint64) == d_c.cpu().numpy().view(np.uint64)).all())
            dig = d_dig.cpu().numpy().view(np.uint64)
            for i in sorted({0, units - 1} | {(units * t) // 16 for t in range(16)}):       # 17 units spread over the batch
                mi, ri = words_to_ints(m_w[i:i + 1])[0], words_to_ints(r_w[i:i + 1])[0]
                _, steps = encrypt_steps(kd["n"], g, mi, ri)
                gs = mi.bit_length() + bin(mi).count("1")
                mine = [(x.q, x.rem) for x in steps[:gs] if x.kind == "mul"] + [(x.q, x.rem) for x in steps[gs:]]
                ok = ok and int(dig[i]) == witness_digest(mine, key.words_out)
            peak_w, _ = imad_peak()
            witness = {"value": world * units / (w_ms * 1e-3), "unit": "units/s", "engine": key.witness_engine, "ms_per_launch": w_ms,
                       "records_per_unit": w_sqr_n + w_mul_n, "mul_mod_per_s": world * units * (w_sqr_n + w_mul_n) / (w_ms * 1e-3),
                       "witness_stream_GBps": world * units * (w_sqr_n + w_mul_n) * 2 * key.words_out * 8 / (w_ms * 1e-3) / 1e9,
                       "chain": {"mod_sqr": w_sqr_n, "mod_mul": w_mul_n, "mac_per_unit": a_wit},
                       "frac_of_imad_peak": units * a_wit / (w_ms * 1e-3) / peak_w,
                       "gpu_launches": int(lib.pb200_kernel_launches() - wl0), "clocks": w_clocks,
                       "ms_each": [a.elapsed_time(b) for a, b in w_evs],
                       "parity": ok,
                       "note": "reference chain (SURVEY.md A.5: bits(n) square_mod + popcount(n) + popcount(m) + 1 mul_mod per unit), exact (q, rem) "
                               "per step folded into a 64-bit digest per unit on the device; ciphertexts equal the fast chain's; the digests of "
                               "17 units spread over the batch are re-derived from the oracle's (q, rem) stream"}

    # ---- K4: advice-cell expansion of mul_mod groups (HBM-bound writer), rank 0 only
    cells = None
    if not args.no_witness and rank == 0:
        cells = cells_leg(key, kd, N_BITS)

    # parity spot check of the last step's output against the CPU port (a checker, never the thing measured)
    parity = None
    if rank == 0:
        from oracle import cpu_ref
        idx = [0, 1, units // 2, units - 1]
        want = cpu_ref.enc_batch(kd["n"], g, N_BITS // 64, m_w[idx], r_w[idx], threads=4)
        got_dev = d_c.cpu().numpy().view(np.uint64)[idx]
        got_host = c_pin.numpy().view(np.uint64)[idx]
        parity = bool((want == got_dev).all() and (want == got_host).all())

    if rank == 0:
        peak, peak_src = imad_peak()
        kernel_s = dev_ms * 1e-3 / args.steps            # one launch per step per rank
        achieved = units * a_enc / kernel_s
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64 columns over signed 28-bit digits (IMAD.WIDE) + s8 x s8 -> s32 (IMMA) for the constant-operand phases", "data": "synthetic",
            "config": {"workload": f"batched encrypt |n|={N_BITS} (4096-bit n^2), {units} units per GPU per step, "
                                   f"{'random g' if args.g == 'rand' else 'g = n+1'}, full-width m and r (BASELINE.json configs[1])",
                       "engine": key.engine, "l2": "flushed between timed steps (256 MiB write)",
                       "chain": {"mod_sqr": n_sqr, "mod_mul": n_mul, "mac_per_enc": a_enc}},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(m_pin.numel() * 8 + r_pin.numel() * 8),
                    "d2h_bytes_per_step": int(c_pin.numel() * 8)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_spot_check": parity,
            "witness": witness,
            "cells": cells,
            "roofline": {"bound": "imad", "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TMAC/s (32x32->64 multiply-accumulate)",
                         "frac": achieved / peak, "traffic": TRAFFIC_BYTES.get((N_BITS, units)), "peak_source": peak_src,
                         "note": "algorithmic MACs = units x (mod_sqr x W_sqr + mod_mul x W_mul), W for 32-bit limbs over n^2 "
                                 "(SURVEY.md 8d), against the measured IMAD.WIDE rate; the block28t engine runs the per-ciphertext "
                                 "products on the IMAD pipe and the two constant-operand Barrett products on the tensor pipe (IMMA), "
                                 "so the fraction can exceed what the IMAD pipe alone could deliver; HBM is idle by design"},
        }
        if not args.no_cpu and world >= 1:
            from oracle import cpu_ref
            threads = cpu_ref.hardware_threads()
            sample = max(64, min(units, int(256 * threads * (2048 / N_BITS) ** 3)))     # ~3-6 s of CPU work at every key size
            v, dt = cpu_baseline(threads, sample, kd)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"first {sample} units of the same batch, {dt:.1f} s, OpenSSL BIGNUM port of src/paillier.rs:87-92"}
            try:        # second line of SURVEY.md 8d: the Python-int restatement on one core, two units
                from oracle.paillier_oracle import paillier_enc_native as _py_enc
                from paillier_halo2_b200.api import words_to_ints as _w2i
                _ms, _rs = _w2i(m_w[:2]), _w2i(r_w[:2])
                _t0 = time.perf_counter()
                for _m, _r in zip(_ms, _rs):
                    _py_enc(kd["n"], g, _m, _r)
                line["cpu_baseline"]["python_int_enc_per_s_one_core"] = 2 / (time.perf_counter() - _t0)
            except Exception as _e:     # a reporting extra must never cost the bench line
                line["cpu_baseline"]["python_int_enc_per_s_one_core"] = None
        emit(line)
    key.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
