"""ncu target at the bench shape: one FRI-commit step of N (default 4096) C2 blobs, device-resident."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import frieda_b200 as F
from bench import synth_blobs, BLOB_LEN
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ctx = F.Context(0)
cfg = F.PcsConfig(4, 0, 20, 20)
L = 1 + ctx.n_inner_layers(BLOB_LEN, cfg)
d_in = torch.from_numpy(synth_blobs(n)).cuda()
d_roots = torch.zeros((n, L, 32), dtype=torch.uint8, device="cuda")
d_last = torch.zeros((n, 1, 4), dtype=torch.int32, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    ctx.fri_commit_batch_ptr(d_in.data_ptr(), BLOB_LEN, BLOB_LEN, n, None, cfg, d_roots.data_ptr(), d_last.data_ptr(), device=True)
torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
print(bytes(d_roots[0, 0].cpu().numpy()).hex())
