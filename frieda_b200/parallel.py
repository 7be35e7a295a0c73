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


def slice_bounds(n_bytes: int, rank: int, world: int) -> Tuple[int, int]:
    """Byte range of the input that `rank` uploads in the sharded-upload mode (16-byte aligned cuts)."""
    per = -(-n_bytes // world)
    per = (per + 15) // 16 * 16
    lo = min(n_bytes, rank * per)
    return lo, min(n_bytes, lo + per)


def commit_split(ctx, data, log_blowup_factor: int, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None, sharded_upload: Optional[bool] = None) -> bytes:
    """commit() of ONE blob with the evaluation domain split across the ranks of `group`.
    Every rank passes the same `data`; every rank returns the same 32-byte root.

    sharded_upload (default: on when world > 1): each rank copies only its 1/world slice of the input
    to its GPU over its own PCIe link and the slices are all-gathered over NVLink (NCCL), instead of
    every rank uploading the whole blob."""
    import numpy as np
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
    if sharded_upload is None:
        sharded_upload = world > 1 and distributed
    if sharded_upload and world > 1:
        host = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), dtype=np.uint8)
        host = host.reshape(-1).view(np.uint8)
        n_bytes = host.size
        per = slice_bounds(n_bytes, 0, world)[1]
        full = torch.empty(per * world, dtype=torch.uint8, device=dev)
        lo, hi = slice_bounds(n_bytes, rank, world)
        mine = full[rank * per: (rank + 1) * per]
        if hi > lo:
            mine[: hi - lo].copy_(torch.from_numpy(host[lo:hi]), non_blocking=True)
        if hi - lo < per:
            mine[hi - lo:].zero_()
        dist.all_gather_into_tensor(full, mine.clone(), group=group)
        torch.cuda.current_stream(dev).synchronize()
        ctx.commit_split_local_device(full.data_ptr(), n_bytes, log_blowup_factor, rank, world, sub.data_ptr())
    else:
        ctx.commit_split_local(data, log_blowup_factor, rank, world, sub.data_ptr())
    if world == 1:
        gathered = sub.reshape(1, 32)
    else:
        gathered = all_gather_roots(sub, group)
        torch.cuda.current_stream(dev).synchronize()
    return ctx.merkle_combine(gathered.data_ptr(), world)
