"""Per-kernel times of one FRI-commit wave (C2 shape) for Merkle tuning probes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frieda_b200 as F
from bench import synth_blobs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = F.Context(0)
blobs = synth_blobs(n)
cfg = F.PcsConfig(4, 0, 20, 20)
for _ in range(2): ctx.fri_commit_batch(blobs, None, cfg)
ctx.set_profiling(True)
roots, last = ctx.fri_commit_batch(blobs, None, cfg)
ctx.set_profiling(False)
prof = ctx.profile_read()
tot = sum(v[1] for v in prof.values())
print(os.environ.get("TAG", ""), f"total {tot:.2f} ms", {k: round(v[1], 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:5]}, roots[0, 0].tobytes().hex()[:16])
