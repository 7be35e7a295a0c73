"""Two real ranks over NVLink (skipped on boxes with one GPU): the split commit through peer-mapped memory
(frieda_commit_split_peers) and through the NCCL form must both give the oracle's root."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import frieda_b200 as F
from frieda_b200 import parallel
from oracle import oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = F.Context(local)
# (8 MiB: slices of 4 MiB are uploaded in parts while the packing kernel of BOTH ranks already waits for them)
for n_bytes, blow in ((100003, 2), (1 << 20, 2), (131072, 4), (3000, 3), (8 << 20, 2)):
    for peers in (False, True):
        # repeated calls reuse the symmetric buffers and advance the barrier epoch; the CONTENT changes every call, so
        # a peer reading a stale copy of another rank's slice or root would give a wrong root
        for k in range(4):
            data = np.frombuffer(O.splitmix64_bytes(0x4652494544414236 + 977 * k + n_bytes, n_bytes), dtype=np.uint8).copy()
            want = O.commit(data.tobytes(), blow)
            assert parallel.commit_split(ctx, data, blow, peer_memory=peers) == want, (n_bytes, blow, peers, k)
assert not parallel._peer_memory_broken
# FRI commit phase of one blob with every layer split over the two ranks: the per-layer exchange of subtree roots as
# NCCL all-gathers (peers False) and inside the library over peer-mapped memory (peers True: one call for all layers;
# run twice so that the root slots are reused under an advanced epoch)
for n_bytes, cfg, seed in ((1 << 20, (2, 0, 20, 12), 7), (131072, (4, 0, 20, 20), None), (8 << 20, (2, 1, 20, 12), 3)):
    data = O.splitmix64_bytes(0x4652494544414236 + n_bytes, n_bytes)
    oroots, olast = O.fri_commit(data, seed, O.make_config(*cfg))
    for peers in (False, True, True):
        roots, last = parallel.fri_commit_split(ctx, data, seed, F.PcsConfig(*cfg), peer_memory=peers)
        assert [r.tobytes() for r in roots] == oroots, (n_bytes, cfg, peers)
        assert [tuple(int(x) for x in q) for q in last] == olast, (n_bytes, cfg, peers)
# ... and the whole proof: replicated grind + queries, owner-serves-path decommitment, shares all-gathered and merged
for n_bytes, cfg, seed in ((1 << 20, (2, 0, 20, 12), 7), (131072, (4, 0, 64, 16), 5)):
    data = O.splitmix64_bytes(0x4652494544414236 + n_bytes, n_bytes)
    oroot, opr = O.prove(data, seed, O.make_config(*cfg))
    for peers in (False, True):
        root, proof = parallel.prove_split(ctx, data, seed, F.PcsConfig(*cfg), peer_memory=peers)
        assert root == oroot and proof.serialize() == opr.serialize(), (n_bytes, cfg, peers)
        assert F.verify_proof(proof, seed)
assert not parallel._peer_memory_broken
ctx.close()
dist.barrier(); dist.destroy_process_group()
if rank == 0: print("MULTI_GPU_OK")
"""


@pytest.mark.gpu
def test_commit_split_two_ranks_peer_memory_and_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
