"""Single-blob latency (C2) and the C5 single-GPU breakdown, with per-kernel event timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import frieda_b200 as F
from bench import synth_blobs
from oracle import oracle as O
ctx = F.Context(0)
cfg = F.PcsConfig(4, 0, 20, 20)
one = torch.from_numpy(synth_blobs(1)).pin_memory().numpy()
for _ in range(5):
    ctx.fri_commit_batch(one, None, cfg)
t0 = time.perf_counter()
for _ in range(50):
    ctx.fri_commit_batch(one, None, cfg)
print("C2 single-blob fri_commit latency ms", (time.perf_counter() - t0) / 50 * 1e3)
ctx.set_profiling(True)
ctx.fri_commit_batch(one, None, cfg)
ctx.set_profiling(False)
pr = ctx.profile_read()
print(" kernels ms", round(sum(v[1] for v in pr.values()), 3), {k: (v[0], round(v[1], 3)) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])})
for _ in range(5):
    ctx.commit(one[0], 4)
t0 = time.perf_counter()
for _ in range(50):
    ctx.commit(one[0], 4)
print("C2 single-blob commit latency ms", (time.perf_counter() - t0) / 50 * 1e3)
for _ in range(3):
    ctx.commit_and_generate_proof(one[0], 1, cfg)
t0 = time.perf_counter()
for _ in range(20):
    ctx.commit_and_generate_proof(one[0], 1, cfg)
print("C2 single-blob prove latency ms", (time.perf_counter() - t0) / 20 * 1e3)
# C5
blob = torch.from_numpy(np.frombuffer(O.splitmix64_bytes(0x4652494544414236, 64 << 20), dtype=np.uint8).copy()).pin_memory().numpy()
for _ in range(2):
    ctx.commit(blob, 2)
t0 = time.perf_counter()
for _ in range(5):
    ctx.commit(blob, 2)
print("C5 commit ms", (time.perf_counter() - t0) / 5 * 1e3)
ctx.set_profiling(True)
ctx.commit(blob, 2)
ctx.set_profiling(False)
pr = ctx.profile_read()
print(" kernels ms", round(sum(v[1] for v in pr.values()), 3), {k: (v[0], round(v[1], 3)) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])})
ctx.set_profiling(True)
ctx.commit_and_generate_proof(one[0], 1, cfg)
ctx.set_profiling(False)
pr = ctx.profile_read()
print("prove kernels ms", round(sum(v[1] for v in pr.values()), 3), {k: (v[0], round(v[1], 3)) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][1])[:8]})
