"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing.

* Independent blobs (BASELINE configs 3/4) shard across ranks with NO data-path collective:
  `shard_blobs` gives each rank a contiguous range; results are gathered by the caller if wanted.
* One oversized blob (config 5): rank r of G = 2^g owns the contiguous bit-reversed index range
  [r N/G, (r+1) N/G) of the evaluation domain = one Merkle subtree at depth g.  Each rank computes
  its subtree root from the whole input (frieda_commit_split_local), the 32-byte roots are
  all-gathered (NCCL over NVLink/NVSwitch when the tensors are CUDA tensors), and every rank hashes
  the top g levels (frieda_merkle_combine).  With peer-mapped symmetric memory (the default on one
  NVLink node) neither exchange is a separate collective: each rank uploads 1/G of the blob, the
  packing kernel reads all G slices in place over NVLink and the combine kernel reads the G roots in
  place (`commit_split_peers`; two stream-ordered barriers, one host synchronisation).  The reference
  has no counterpart (it is a single-threaded CPU path, src/commit.rs:11-23); the result is the same root.
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


class _PeerState:
    """Peer-mapped (symmetric) buffers of one process group on one context: every rank's input slice and
    every rank's subtree root are readable from every GPU over NVLink / NVSwitch."""

    def __init__(self, dev, group, slice_cap: int):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        g = group if group is not None else dist.group.WORLD
        self.cap = slice_cap
        self.buf = symm_mem.empty(slice_cap, dtype=torch.uint8, device=dev)
        self.h_in = symm_mem.rendezvous(self.buf, group=g)
        self.roots = symm_mem.empty(64 * 32, dtype=torch.uint8, device=dev)
        self.h_roots = symm_mem.rendezvous(self.roots, group=g)
        # signal words of the library's own stream-ordered barriers and upload flags (8 channels x 64 ranks), zero
        # before first use
        self.flags = symm_mem.empty(512, dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.h_flags = symm_mem.rendezvous(self.flags, group=g)
        # per-layer subtree roots of a split FRI commit (frieda_fri_split_layers_peers): slot l of rank r's area
        self.fri_roots = symm_mem.empty(64 * 32, dtype=torch.uint8, device=dev)
        self.h_fri_roots = symm_mem.rendezvous(self.fri_roots, group=g)
        self.h_flags.barrier(channel=0)
        torch.cuda.current_stream(dev).synchronize()
        self.epoch = 0
        # the pointer lists handed to the library on every call (the properties rebuild them on each access)
        self.slice_ptrs = [int(p) for p in self.h_in.buffer_ptrs]
        self.root_ptrs = [int(p) + 32 * r for r, p in enumerate(self.h_roots.buffer_ptrs)]
        self.flag_ptrs = [int(p) for p in self.h_flags.buffer_ptrs]
        self.fri_root_ptrs = [int(p) for p in self.h_fri_roots.buffer_ptrs]


_peer_states = {}
_peer_memory_broken = False


def release_peer_memory(ctx=None):
    """Frees the symmetric buffers (of one context, or all).  Called by Context.close() and at interpreter exit,
    before the streams they were used on disappear."""
    for key in [k for k in _peer_states if ctx is None or k[0] == ctx.device]:
        del _peer_states[key]


import atexit  # noqa: E402

atexit.register(release_peer_memory)


def _peer_state(ctx, group, slice_len: int):
    """Collective (every rank of the group calls it with the same slice_len); grow-only."""
    import torch
    key = (ctx.device, id(group))
    st = _peer_states.get(key)
    if st is None or st.cap < slice_len:
        cap = 1 << 20
        while cap < slice_len:
            cap <<= 1
        st = _PeerState(torch.device("cuda", ctx.device), group, cap)
        _peer_states[key] = st
    return st


def commit_split_peers(ctx, host, n_bytes: int, log_blowup_factor: int, rank: int, world: int, group=None,
                       st=None, resident: bool = False) -> bytes:
    """The split commit over peer memory: ONE library call, ordered on the context's stream with one host
    synchronisation at the end (frieda_commit_split_peers): upload my slice -> barrier kernel -> pack straight
    out of the peers' slices (the all-gather is the packing kernel's loads over NVLink) -> LDE + Merkle of my
    subtree -> root into my symmetric slot -> barrier kernel -> combine kernel reads the peers' roots in place.
    No NCCL call on the data path."""
    per = slice_bounds(n_bytes, 0, world)[1]
    if st is None:
        st = _peer_state(ctx, group, per)
    st.epoch += 1
    # resident: the slices of a previous call with the same blob are still in the symmetric buffers (inputs in HBM)
    return ctx.commit_split_peers(host, log_blowup_factor, rank, st.slice_ptrs, per, st.root_ptrs, st.flag_ptrs, st.epoch,
                                  resident_len=n_bytes if resident else None)


def prove_split(ctx, data, seed, cfg, group=None, rank: Optional[int] = None, world: Optional[int] = None,
                all_gather=None, all_gather_bytes=None, peer_memory: Optional[bool] = None):
    """commit_and_generate_proof (src/proof.rs:32-77) of ONE blob split across the ranks of `group`: the split FRI commit
    with every tree kept, then on every rank proof of work + queries (replicated) and the rank's share of the
    decommitment; the shares are all-gathered (a few hundred KB) and merged on every rank.  Returns (root, Proof),
    byte-identical to Context.commit_and_generate_proof on one GPU.

    all_gather_bytes(b) -> list of `world` byte strings in rank order; default torch.distributed.all_gather_object."""
    import torch.distributed as dist
    from .api import split_assemble
    distributed = dist.is_available() and dist.is_initialized()
    if world is None:
        world = dist.get_world_size(group) if distributed else 1
    roots, _ = fri_commit_split(ctx, data, seed, cfg, group, rank, world, all_gather, keep_trees=True,
                                peer_memory=peer_memory)
    share = ctx.fri_split_decommit()
    if all_gather_bytes is None:
        def all_gather_bytes(b):
            if world == 1:
                return [b]
            if not getattr(ctx, "is_cuda", True):
                out = [None] * world
                dist.all_gather_object(out, b, group=group)
                return out
            # two tensor collectives (lengths, then the shares padded to the longest) instead of all_gather_object's
            # pickling round trips: the shares are a few hundred KB
            import numpy as np
            import torch
            dev = torch.device("cuda", ctx.device)
            lens = torch.zeros(world, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(lens, torch.tensor([len(b)], dtype=torch.int64, device=dev), group=group)
            lens = [int(x) for x in lens.cpu()]
            cap = (max(lens) + 15) // 16 * 16
            mine = torch.zeros(cap, dtype=torch.uint8, device=dev)
            mine[: len(b)].copy_(torch.from_numpy(np.frombuffer(b, dtype=np.uint8).copy()))
            allb = torch.empty(world * cap, dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(allb, mine, group=group)
            host = allb.cpu().numpy()
            return [host[r * cap: r * cap + lens[r]].tobytes() for r in range(world)]
    return roots[0].tobytes(), split_assemble(all_gather_bytes(share))


def fri_commit_split(ctx, data, seed, cfg, group=None, rank: Optional[int] = None, world: Optional[int] = None,
                     all_gather=None, keep_trees: bool = False, peer_memory: Optional[bool] = None):
    """FriProver::commit (src/proof.rs:52-57) of ONE blob with every layer split across the ranks of `group`
    (frieda_fri_split_*): per split layer a rank-local fused fold + subtree, an all-gather of `world` 32-byte
    subtree roots (NCCL over NVLink), the top levels and the channel step on every rank; then an all-gather of the
    first unsplit layer's columns and the ordinary kernels.  Every rank passes the same data / seed / cfg and returns
    the same (layer_roots (n_layers, 32) uint8, last_layer_poly (2^log_last, 4) uint32), bit-identical to
    Context.fri_commit_batch on one GPU.

    all_gather(tensor) -> tensor of shape (world, *tensor.shape): the transport; default torch.distributed over
    `group` (tests with several contexts on one GPU pass their own).

    peer_memory (default: tried first on real ranks): the split layers run as ONE library call per rank
    (frieda_fri_split_layers_peers) whose per-layer exchange of subtree roots is done by the library's kernels over
    peer-mapped symmetric memory -- no host-driven collective per layer; otherwise one NCCL all-gather of `world`
    roots per layer."""
    import torch
    import torch.distributed as dist
    global _peer_memory_broken
    distributed = dist.is_available() and dist.is_initialized()
    if world is None:
        world = dist.get_world_size(group) if distributed else 1
    if rank is None:
        rank = dist.get_rank(group) if distributed else 0
    if not is_pow2(world):
        raise ValueError("fri_commit_split needs a power-of-two number of ranks")
    import contextlib
    on_gpu = getattr(ctx, "is_cuda", True)  # (the CPU tests of this sequencing drive a stand-in context over gloo)
    dev = torch.device("cuda", ctx.device) if on_gpu else torch.device("cpu")
    # the collectives are ordered on the context's stream, like its kernels
    scope = torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream_ptr, device=dev)) if on_gpu else contextlib.nullcontext()
    real_ranks = on_gpu and world > 1 and distributed and all_gather is None
    if all_gather is None:
        def all_gather(t):
            if world == 1:
                return t.reshape((1,) + tuple(t.shape))
            out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(out, t.reshape((1,) + tuple(t.shape)).contiguous(), group=group)
            return out
    st = None
    host = per = None
    if real_ranks:
        import numpy as np
        host = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), dtype=np.uint8)
        host = host.reshape(-1).view(np.uint8)
        per = slice_bounds(host.size, 0, world)[1]
    if real_ranks and peer_memory is not False and (peer_memory or not _peer_memory_broken):
        try:
            st = _peer_state(ctx, group, max(per, 1 << 20))
        except (ImportError, RuntimeError, AttributeError) as e:
            # collective: fails on every rank alike, so all of them take the NCCL form below
            if peer_memory:
                raise
            _peer_memory_broken = True
            import warnings
            warnings.warn(f"peer-mapped split FRI commit unavailable ({e}); using NCCL all-gathers")
    full = None
    shape = None
    if st is not None:
        # the blob travels like commit_split_peers' input: each rank uploads its slice into the symmetric buffer and
        # the packing kernel reads all slices where they lie
        st.epoch += 1
        shape = ctx.fri_split_begin_peers(host, seed, cfg, rank, st.slice_ptrs, per, st.flag_ptrs, st.epoch,
                                          keep_trees=keep_trees)
    elif real_ranks and per >= (1 << 16):
        # every rank needs the whole blob: each uploads 1/world of it over its own PCIe link, NCCL all-gathers the
        # slices over NVLink (as commit_split's NCCL form does), and the library starts from device memory
        with scope:
            full = torch.empty(per * world, dtype=torch.uint8, device=dev)
            lo, hi = slice_bounds(host.size, rank, world)
            mine = torch.zeros(per, dtype=torch.uint8, device=dev)
            if hi > lo:
                mine[: hi - lo].copy_(torch.from_numpy(host[lo:hi]), non_blocking=True)
            dist.all_gather_into_tensor(full, mine, group=group)
        shape = ctx.fri_split_begin(None, seed, cfg, rank, world, device_ptr=full.data_ptr(), length=host.size,
                                    keep_trees=keep_trees)
    if shape is None:
        shape = ctx.fri_split_begin(data, seed, cfg, rank, world, keep_trees=keep_trees)
    n_split, n_layers, handoff_log = shape
    with scope:
        # (allocated inside the scope: every tensor the library writes is touched on the context's stream only)
        if st is not None:
            ctx.fri_split_layers_peers(st.fri_root_ptrs, st.flag_ptrs, st.epoch)
        else:
            sub = torch.empty(32, dtype=torch.uint8, device=dev)
            for layer in range(n_split):
                ctx.fri_split_layer(layer, sub.data_ptr())
                roots = all_gather(sub)
                ctx.fri_split_combine(layer, roots.data_ptr())
        mine = torch.empty(4 << handoff_log, dtype=torch.int32, device=dev)
        ctx.fri_split_handoff(mine.data_ptr())
        cols_all = all_gather(mine)
        out = ctx.fri_split_finish(cols_all.data_ptr(), n_layers, cfg.log_last_layer_degree_bound)
    return out


def commit_split(ctx, data, log_blowup_factor: int, group=None, rank: Optional[int] = None,
                 world: Optional[int] = None, sharded_upload: Optional[bool] = None,
                 peer_memory: Optional[bool] = None) -> bytes:
    """commit() of ONE blob with the evaluation domain split across the ranks of `group`.
    Every rank passes the same `data`; every rank returns the same 32-byte root.

    sharded_upload (default: on when world > 1): each rank copies only its 1/world slice of the input
    to its GPU over its own PCIe link instead of every rank uploading the whole blob.
    peer_memory (default: tried first when sharded_upload is on): the slices and the subtree roots live in
    peer-mapped symmetric memory and are read in place over NVLink by the packing and combine kernels
    (`commit_split_peers`); otherwise the slices and the roots are all-gathered with NCCL."""
    import numpy as np
    import torch
    import torch.distributed as dist
    global _peer_memory_broken
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
    if sharded_upload and world > 1 and peer_memory is not False and (peer_memory or not _peer_memory_broken):
        host = data if isinstance(data, np.ndarray) else np.frombuffer(bytes(data), dtype=np.uint8)
        host = host.reshape(-1).view(np.uint8)
        st = None
        try:
            st = _peer_state(ctx, group, slice_bounds(host.size, 0, world)[1])
        except (ImportError, RuntimeError, AttributeError) as e:
            # symmetric memory unavailable (old torch, no P2P mapping): the allocation / rendezvous is collective
            # and fails on every rank alike, so all of them take the NCCL path below
            if peer_memory:
                raise
            _peer_memory_broken = True
            import warnings
            warnings.warn(f"peer-mapped split commit unavailable ({e}); using NCCL all-gathers")
        if st is not None:
            return commit_split_peers(ctx, host, host.size, log_blowup_factor, rank, world, group, st)
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
