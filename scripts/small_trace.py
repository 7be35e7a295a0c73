"""FRIEDA_SMALL_TRACE=1 python scripts/small_trace.py: per-layer stage times of the latency path (CTA 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frieda_b200 as F
from bench import synth_blobs
ctx = F.Context(0)
one = synth_blobs(1)
cfg = F.PcsConfig(4, 0, 20, 20)
for _ in range(3):
    ctx.fri_commit_batch(one, None, cfg)
