import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import frieda_b200 as F
ctx = F.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream_ptr)
nb, p, beta, n_felts = 256, 14, 4, 34953
D = p + beta
rng = np.random.default_rng(0)
coef = np.zeros((nb, 4 << p), dtype=np.uint32)
coef[:, :n_felts] = rng.integers(0, (1 << 31) - 1, (nb, n_felts), dtype=np.uint32)
d_coef = torch.from_numpy(coef.view(np.int32)).cuda()
d_eval = torch.empty((nb, 4 << D), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
f = lambda: ctx.pass_lde(d_coef.data_ptr(), p, beta, nb, n_felts, d_eval.data_ptr())
for _ in range(3): f()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream.synchronize(); e0.record(stream)
for _ in range(20): f()
e1.record(stream); stream.synchronize()
ms = e0.elapsed_time(e1) / 20
b = nb * (16 * (1 << p) + 16 * (1 << D))
print(f"lde {ms:.4f} ms  {b/ms/1e6:.0f} GB/s  {b/ms/1e6/6552:.3f}", "checksum", int(d_eval.view(torch.int64).sum().item()) & 0xffffffff)
