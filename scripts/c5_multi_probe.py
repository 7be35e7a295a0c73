"""torchrun target: the split commit of one 64 MiB blob over peer memory vs the NCCL all-gather path."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import frieda_b200 as F
from frieda_b200.parallel import commit_split
from oracle import oracle as O
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = F.Context(local)
C5_ROOT = "7ab1d35e23dcb4b678524912e0fc0cdbcb6efaa34aa9f480f97f3138b3f0ee07"
blob = torch.frombuffer(bytearray(O.splitmix64_bytes(0x4652494544414236, 64 << 20)), dtype=torch.uint8).pin_memory().numpy()
small = np.frombuffer(O.splitmix64_bytes(0x4652494544414236, 100003), dtype=np.uint8).copy()
want_small = O.commit(small.tobytes(), 2).hex()
for peers in (False, True):
    assert commit_split(ctx, small, 2, peer_memory=peers).hex() == want_small, ("small", peers)
    root = commit_split(ctx, blob, 2, peer_memory=peers)
    assert root.hex() == C5_ROOT, (peers, root.hex())
    for _ in range(3): commit_split(ctx, blob, 2, peer_memory=peers)
    ts = []
    for _ in range(10):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        commit_split(ctx, blob, 2, peer_memory=peers)
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        ts.append(float(dt.item()))
    if rank == 0:
        print(f"world {world} peer_memory={peers}: roots ok; ms per blob min {min(ts)*1e3:.3f} median {sorted(ts)[5]*1e3:.3f}", flush=True)
    # one profiled call per rank: CUDA-event time of every kernel on the context's stream (barrier kernels = waiting)
    dist.barrier(); torch.cuda.synchronize()
    ctx.profile_read(reset=True); ctx.set_profiling(True)
    t0 = time.perf_counter()
    commit_split(ctx, blob, 2, peer_memory=peers)
    wall = time.perf_counter() - t0
    ctx.set_profiling(False)
    prof = ctx.profile_read(reset=True)
    for r in range(world):
        dist.barrier()
        if r == rank and r in (0, world - 1):
            print(f"  rank {rank} peers={peers} wall {wall*1e3:.3f} ms kernels {sum(v[1] for v in prof.values()):.3f} ms",
                  {k: (v[0], round(v[1], 3)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}, flush=True)
dist.barrier(); dist.destroy_process_group()
