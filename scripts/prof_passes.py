"""ncu target: the standalone LDE and fold passes at the bench's shape (256 C2 blobs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import frieda_b200 as F
ctx = F.Context(0)
nb, p, beta, n_felts = 256, 14, 4, 34953
D = p + beta
rng = np.random.default_rng(0)
coef = np.zeros((nb, 4 << p), dtype=np.uint32)
coef[:, :n_felts] = rng.integers(0, (1 << 31) - 1, (nb, n_felts), dtype=np.uint32)
d_coef = torch.from_numpy(coef.view(np.int32)).cuda()
d_eval = torch.empty((nb, 4 << D), dtype=torch.int32, device="cuda")
d_line = torch.empty((nb, 4 << (D - 1)), dtype=torch.int32, device="cuda")
d_line2 = torch.empty((nb, 4 << (D - 2)), dtype=torch.int32, device="cuda")
d_alpha = torch.from_numpy(rng.integers(0, (1 << 31) - 1, (nb, 4), dtype=np.uint32).view(np.int32)).cuda()
torch.cuda.synchronize()
for _ in range(2):
    ctx.pass_lde(d_coef.data_ptr(), p, beta, nb, n_felts, d_eval.data_ptr())
    ctx.pass_fold(d_eval.data_ptr(), D, True, nb, d_alpha.data_ptr(), d_line.data_ptr())
    ctx.pass_fold(d_line.data_ptr(), D - 1, False, nb, d_alpha.data_ptr(), d_line2.data_ptr())
torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
print("done")
