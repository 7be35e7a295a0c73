import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import frieda_b200 as F
from bench import synth_blobs
ctx = F.Context(0)
cfg = F.PcsConfig(4, 0, 20, 20)
one = torch.from_numpy(synth_blobs(1)).pin_memory().numpy()
for _ in range(3):
    ctx.commit_and_generate_proof(one[0], 1, cfg)
os.environ["FRIEDA_TRACE"] = "1"
t0 = time.perf_counter()
ctx.commit_and_generate_proof(one[0], 1, cfg)
print("wall ms", (time.perf_counter() - t0) * 1e3)
