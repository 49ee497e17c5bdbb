"""Multi-GPU sharding of the Paillier hot path.

Units are independent, so encrypt / add / witness shard by contiguous index range with NO data-path
collective.  The tally has one real exchange step: every rank folds its shard to one partial product on
its own GPU, the G partials (G x 2|n|/8 bytes) are all-gathered (NCCL over NVLink on GPUs, gloo in the CPU
tests) and combined with G-1 modular multiplications.  The product is commutative and associative, so the
result is bit-identical for any G and any shard shape.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def shard_range(count: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous partition [i*N/G, (i+1)*N/G) of `count` units over `world` ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return count * rank // world, count * (rank + 1) // world


def all_gather_words(partial, world: int, group=None):
    """all-gather one (words,) int64 tensor per rank -> (world, words) tensor on the same device."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return partial.reshape(1, -1)
    out = torch.empty(world * partial.numel(), dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous().reshape(-1), group=group)
    return out.reshape(world, partial.numel())


def tally_sharded(local_tally: Callable, combine: Callable, partial_buffer, world: int, group=None):
    """local_tally() must leave this rank's partial in `partial_buffer`; returns combine(all partials)."""
    local_tally()
    gathered = all_gather_words(partial_buffer, world, group)
    return combine(gathered)


def exchange_handles(handle: bytes, world: int, group=None) -> List[bytes]:
    """all-gather of the ranks' 64-byte mailbox handles (host side, once per key)."""
    import torch.distributed as dist

    if world == 1:
        return [handle]
    out: List[bytes] = [b""] * world
    dist.all_gather_object(out, handle, group=group)
    return out


def connect_tally_peers(key, rank: int, world: int, group=None) -> str:
    """One-off setup of the fused multi-GPU tally: every rank exports its mailbox handle, the handles are all-gathered on the
    host and every rank maps its peers' mailboxes (CUDA IPC, NVLink peer access).  Returns "peer-memory", or "nccl" when some
    rank could not map its peers (then tally_sharded_gpu falls back to an all-gather of the partials)."""
    import torch
    import torch.distributed as dist

    ok = 1
    try:
        handles = exchange_handles(key.tally_peer_export(), world, group)
        key.tally_peer_connect(rank, world, handles)
    except Exception:       # the collective decision below must be reached by every rank
        ok = 0
    if world > 1:
        t = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{key.device}" if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        ok = int(t.item())
    return "peer-memory" if ok else "nccl"


def tally_sharded_gpu(key, d_c, count: int, world: int, group=None, exchange: str = "nccl", out=None, sync: bool = True):
    """Device path: d_c = this rank's shard (torch int64 tensor on the key's device, count x words_out).
    Returns a (words_out,) int64 device tensor holding the full product mod n^2 (same on every rank).
    exchange = "peer-memory" (after connect_tally_peers): ONE kernel per GPU, partials cross NVLink inside it;
    exchange = "nccl": per-GPU fold, all-gather of the partials, combine fold."""
    import torch

    if out is None:
        out = torch.empty(key.words_out, dtype=torch.int64, device=d_c.device)
    if exchange == "peer-memory":
        key.tally_peer_dev(d_c.data_ptr(), count, out.data_ptr())
        if sync:
            key.sync()
        return out
    partial = torch.empty(key.words_out, dtype=torch.int64, device=d_c.device)

    def local():
        key.tally_dev(d_c.data_ptr(), count, partial.data_ptr())
        key.sync()

    def combine(g):
        g = g.contiguous()
        key.tally_dev(g.data_ptr(), g.shape[0], out.data_ptr())
        key.sync()
        return out

    return tally_sharded(local, combine, partial, world, group)
