"""CPU tests of the N>1 host logic: index sharding and the all-gather + combine of the tally, run with
world_size 2 over gloo.  The per-rank partial products come from the oracle here (no GPU in this suite);
on the GPU box tests/test_gpu_parity.py::test_tally_sharded_single_gpu drives the same code with the device path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.paillier_oracle import tally_native
from paillier_halo2_b200 import workload
from paillier_halo2_b200.api import ints_to_words, words_to_ints
from paillier_halo2_b200.shard import shard_range, tally_sharded


def test_shard_range_partitions_exactly():
    for count in (0, 1, 7, 64, 65537, 1 << 20):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(count, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == count
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_hex, cs_hex, want_hex, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = int(n_hex, 16)
    cs = [int(c, 16) for c in cs_hex]
    wo = (2 * n.bit_length() + 63) // 64
    lo, hi = shard_range(len(cs), rank, world)
    partial = torch.zeros(wo, dtype=torch.int64)

    def local():
        v = tally_native(n, cs[lo:hi])
        partial.copy_(torch.from_numpy(ints_to_words([v], wo)[0].view("int64")))

    def combine(g):
        return tally_native(n, words_to_ints(g.numpy().view("uint64")))

    got = tally_sharded(local, combine, partial, world)
    q.put((rank, got == int(want_hex, 16)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_tally_allgather_combine_gloo(world):
    kd = workload.load_key(256)
    n = kd["n"]
    cs = words_to_ints(workload.ciphertexts(256, 101, n))
    want = tally_native(n, cs)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, hex(n), [hex(c) for c in cs], hex(want), q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]


class _FakeKey:
    """Stands in for PaillierKey in the handle exchange (no GPU in this suite): records what connect receives."""

    def __init__(self, rank, fail_on=None):
        self.rank, self.fail_on, self.device, self.connected = rank, fail_on, 0, None

    def tally_peer_export(self):
        return bytes([self.rank]) * 64

    def tally_peer_connect(self, rank, world, handles):
        if self.fail_on == rank:
            raise RuntimeError("cudaIpcOpenMemHandle failed")
        self.connected = (rank, world, list(handles))


def _peer_worker(rank, world, port, fail_on, q):
    from paillier_halo2_b200.shard import connect_tally_peers

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    key = _FakeKey(rank, fail_on)
    mode = connect_tally_peers(key, rank, world)
    ok = mode == ("nccl" if fail_on is not None else "peer-memory")
    if key.connected is not None:
        r, w, hs = key.connected
        ok = ok and r == rank and w == world and hs == [bytes([i]) * 64 for i in range(world)]
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("fail_on", [None, 1])
def test_peer_handle_exchange_gloo(fail_on):
    """connect_tally_peers: handles all-gathered in rank order; if ANY rank cannot map its peers, EVERY rank falls back."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, fail_on, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
