"""Times the proof path on N C2-shaped blobs and prints wall vs per-kernel time."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import frieda_b200 as F
from bench import synth_blobs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = F.Context(0)
blobs = torch.from_numpy(synth_blobs(n)).pin_memory().numpy()
cfg = F.PcsConfig(4, 0, 64, 20)
seeds = list(range(n))
ctx.prove_batch(blobs, seeds, cfg)  # (result dropped at once: the timed call reuses the freed memory)
ctx.set_profiling(True)
t0 = time.perf_counter()
roots, proofs = ctx.prove_batch(blobs, seeds, cfg)
dt = time.perf_counter() - t0
ctx.set_profiling(False)
prof = ctx.profile_read()
print(f"n={n} wall {dt*1e3:.1f} ms -> {n/dt:.0f} proofs/s; kernels {sum(v[1] for v in prof.values()):.1f} ms")
print({k: round(v[1], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:6]})
print("mean nonce", sum(p.proof_of_work for p in proofs) / n)
