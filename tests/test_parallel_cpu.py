"""Host-side multi-GPU logic on CPU: blob sharding, and the subtree-root exchange of the split-blob
path over a world_size-2 gloo group (the CUDA ranks use the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest

from frieda_b200 import parallel
from oracle import oracle as O


def test_shard_blobs_partitions_exactly():
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_blobs(n, r, world) for r in range(world)]
            assert spans[0][0] == 0
            assert sum(c for _, c in spans) == n
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_blobs(4, 4, 4)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, level_nodes, root_hex, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # each rank contributes "its" subtree root (precomputed by the oracle for this test)
        sub = torch.from_numpy(np.frombuffer(level_nodes[rank], dtype=np.uint8).copy())
        gathered = parallel.all_gather_roots(sub)
        got = gathered.numpy().tobytes()
        ok = got == b"".join(level_nodes)
        # top levels over the gathered roots with the oracle's compression (checker side)
        level = [np.frombuffer(x, dtype="<u4") for x in level_nodes]
        while len(level) > 1:
            level = [np.array(O.blake2s_compress([0] * 8, [int(v) for v in level[2 * i]] + [int(v) for v in
                                                                                            level[2 * i + 1]]),
                              dtype="<u4") for i in range(len(level) // 2)]
        ok = ok and level[0].tobytes().hex() == root_hex
        start, count = parallel.shard_blobs(4096, rank, world)
        ok = ok and (start, count) == (rank * 2048, 2048)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_subtree_root_exchange_gloo_world2():
    import torch.multiprocessing as mp
    data = bytes(i % 256 for i in range(20000))
    t = O.trace(data, None, O.make_config(2, 0, 4, 1), stop_after_fri=True, with_trees=True)
    world = 2
    nodes = [t.tree_levels[0][1][r].tobytes() for r in range(world)]
    root_hex = t.tree_levels[0][0].tobytes().hex()
    assert root_hex == O.commit(data, 2).hex()
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, world, port, nodes, root_hex, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_slice_bounds_cover_the_blob():
    for n in (0, 1, 15, 16, 17, 1000, 64 << 20, (64 << 20) + 5):
        for world in (1, 2, 4, 8):
            spans = [parallel.slice_bounds(n, r, world) for r in range(world)]
            per = spans[0][1] - spans[0][0] if n else 0
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, _) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(lo % 16 == 0 or lo == n for lo, _ in spans)
            assert all(lo == min(n, r * per) for r, (lo, _) in enumerate(spans)) or n == 0


# ------------------------------------------------------------------ split-blob FRI: the driver's sequencing over gloo
class _OracleSplitCtx:
    """Stand-in for a CUDA context behind parallel.fri_commit_split: every frieda_fri_split_* step answered from the
    oracle's trace of the whole blob, restricted to this rank's index range -- and every INPUT the driver hands back
    (gathered subtree roots, gathered columns) checked against that trace.  Pointers are host addresses here."""
    is_cuda = False
    device = 0
    MIN_LOG = 10  # csrc/ctx.cu: SPLIT_MIN_LOG

    def fri_split_begin(self, data, seed, cfg, rank, world, keep_trees=False):
        self.t = O.trace(bytes(data), seed, O.make_config(cfg.log_blowup_factor, cfg.log_last_layer_degree_bound,
                                                          cfg.n_queries, cfg.pow_bits), stop_after_fri=True, with_trees=True)
        self.rank, self.world, self.gl = rank, world, world.bit_length() - 1
        logs = self.t.layer_logs
        self.n_split = sum(1 for lg in logs if lg >= self.gl + self.MIN_LOG)
        self.combined = []
        nxt = logs[self.n_split] if self.n_split < len(logs) else self.t.last_eval.shape[1].bit_length() - 1
        return self.n_split, len(logs), nxt - self.gl

    def fri_split_layer(self, layer, ptr):
        import ctypes
        node = self.t.tree_levels[layer][self.gl][self.rank].tobytes()   # root of my subtree = node (level g, index rank)
        ctypes.memmove(ptr, node, 32)

    def fri_split_combine(self, layer, ptr):
        import ctypes
        got = ctypes.string_at(ptr, 32 * self.world)
        assert got == self.t.tree_levels[layer][self.gl].tobytes(), f"layer {layer}: gathered subtree roots differ"
        assert layer == len(self.combined)
        self.combined.append(layer)

    def _next_cols(self):
        s = self.n_split
        return self.t.layer_columns[s] if s < len(self.t.layer_logs) else self.t.last_eval

    def fri_split_handoff(self, ptr):
        import ctypes
        assert self.combined == list(range(self.n_split))
        cols = self._next_cols()
        m = cols.shape[1] // self.world
        mine = np.ascontiguousarray(cols[:, self.rank * m:(self.rank + 1) * m])
        ctypes.memmove(ptr, mine.ctypes.data, mine.nbytes)

    def fri_split_finish(self, ptr, n_layers, log_last):
        import ctypes
        cols = self._next_cols()
        m = cols.shape[1] // self.world
        got = np.frombuffer(ctypes.string_at(ptr, cols.nbytes), dtype=np.uint32).reshape(self.world, 4, m)
        want = np.stack([cols[:, r * m:(r + 1) * m] for r in range(self.world)])
        assert np.array_equal(got, want), "gathered columns are not [rank][column][local index]"
        roots = np.stack([lv[0].reshape(32) for lv in self.t.tree_levels])
        return roots, np.array(self.t.last_layer_poly, dtype=np.uint32)


def _split_worker(rank, world, port, q):
    import torch.distributed as dist
    import frieda_b200 as F
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ok = True
        for n, cfg, seed in ((40000, (3, 1, 8, 4), 5), (131072, (2, 0, 8, 4), None)):
            data = bytes((i * 7 + n) % 256 for i in range(n))
            roots, last = parallel.fri_commit_split(_OracleSplitCtx(), data, seed, F.PcsConfig(*cfg))
            oroots, olast = O.fri_commit(data, seed, O.make_config(*cfg))
            ok = ok and [r.tobytes() for r in roots] == oroots and [tuple(int(x) for x in v) for v in last] == olast
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_fri_commit_split_sequencing_gloo_world2():
    import torch.multiprocessing as mp
    world = 2
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_split_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
