"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

* Independent blobs (BASELINE configs 3/4) shard across ranks with NO data-path collective:
  `shard_blobs` gives each rank a contiguous range; results are gathered by the caller if wanted.
* One oversized blob (config 5): rank r of G = 2^g owns the contiguous bit-reversed index range
  [r N/G, (r+1) N/G) of the evaluation domain = one Merkle subtree at depth g.  Each rank computes
  its subtree root from the whole input (frieda_commit_split_local), the 32-byte roots are
  all-gathered (NCCL over NVLink/NVSwitch when the tensors are CUDA tensors), and every rank hashes
  the top g levels (frieda_merkle_combine).  The reference has no counterpart (it is a
  single-threaded CPU path, src/commit.rs:11-23); the result is the same root.
"""
from __future__ import annotations

from typing import Optional, Tuple


def shard_blobs(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, start + count) of n blobs for `rank`; remainder spread over the first ranks."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def is_pow2(x: int) -> bool:
    return x > 0 and (x & (x - 1)) == 0


def all_gather_roots(subroot, group=None):
    """All-gathers one 32-byte root per rank into a (world, 32) uint8 tensor in rank order.
    Works on CUDA tensors (NCCL) and CPU tensors (gloo)."""
    import torch
    import torch.distributed as dist
    assert subroot.dtype == torch.uint8 and subroot.numel() == 32
    world = dist.get_world_size(group)
    out = torch.empty((world, 32), dtype=torch.uint8, device=subroot.device)
    dist.all_gather_into_tensor(out, subroot.reshape(1, 32).contiguous(), group=group)
    return out


def commit_split(ctx, data, log_blowup_factor: int, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None) -> bytes:
    """commit() of ONE blob with the evaluation domain split across the ranks of `group`.
    Every rank passes the same `data`; every rank returns the same 32-byte root."""
    import torch
    import torch.distributed as dist
    distributed = dist.is_available() and dist.is_initialized()
    if world is None:
        world = dist.get_world_size(group) if distributed else 1
    if rank is None:
        rank = dist.get_rank(group) if distributed else 0
    if not is_pow2(world):
        raise ValueError("commit_split needs a power-of-two number of ranks")
    dev = torch.device("cuda", ctx.device)
    sub = torch.zeros(32, dtype=torch.uint8, device=dev)
    ctx.commit_split_local(data, log_blowup_factor, rank, world, sub.data_ptr())
    if world == 1:
        gathered = sub.reshape(1, 32)
    else:
        gathered = all_gather_roots(sub, group)
        torch.cuda.current_stream(dev).synchronize()
    return ctx.merkle_combine(gathered.data_ptr(), world)
