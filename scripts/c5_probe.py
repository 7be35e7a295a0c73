import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import frieda_b200 as F
from oracle import oracle as O
ctx = F.Context(0)
blob = torch.from_numpy(np.frombuffer(O.splitmix64_bytes(0x4652494544414236, 64 << 20), dtype=np.uint8).copy()).pin_memory().numpy()
for _ in range(3):
    r = ctx.commit(blob, 2)
ts = []
for _ in range(10):
    t0 = time.perf_counter(); ctx.commit(blob, 2); ts.append((time.perf_counter() - t0) * 1e3)
print("chunked" if not os.environ.get("FRIEDA_NO_CHUNKED") else "single-copy", "C5 commit ms min/med", round(min(ts), 3), round(sorted(ts)[5], 3), r.hex()[:16])
# raw H2D time of 64 MiB for reference
d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
h = torch.from_numpy(blob)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); print("H2D 64MiB ms", (time.perf_counter() - t0) / 5 * 1e3)
