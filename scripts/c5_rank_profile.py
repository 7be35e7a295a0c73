"""Kernel times of ONE rank's share of the split commit (64 MiB blob, blowup 2^2) for world = 1, 2, 4, 8 on one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import frieda_b200 as F
from frieda_b200.parallel import slice_bounds
from oracle import oracle as O
ctx = F.Context(0)
n = 64 << 20
data = torch.frombuffer(bytearray(O.splitmix64_bytes(0x4652494544414236, n)), dtype=torch.uint8).cuda()
stream = torch.cuda.ExternalStream(ctx.stream_ptr)
for world in (1, 2, 4, 8):
    per = slice_bounds(n, 0, world)[1]
    slices = [data[r * per: (r + 1) * per] for r in range(world)]
    root = torch.zeros(32, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.commit_split_local_peers([s.data_ptr() for s in slices], per, n, 2, 0, root.data_ptr())
    stream.synchronize()
    ctx.profile_read(reset=True); ctx.set_profiling(True)
    t0 = time.perf_counter()
    ctx.commit_split_local_peers([s.data_ptr() for s in slices], per, n, 2, 0, root.data_ptr())
    stream.synchronize()
    dt = time.perf_counter() - t0
    ctx.set_profiling(False)
    prof = ctx.profile_read(reset=True)
    print(f"world {world}: wall {dt*1e3:.3f} ms, kernels {sum(v[1] for v in prof.values()):.3f} ms",
          {k: (v[0], round(v[1], 3)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})
