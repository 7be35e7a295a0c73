"""Small target for `ncu --set full`: one FRI-commit wave of N blobs (C2 shape)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import frieda_b200 as F  # noqa: E402
from bench import synth_blobs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
ctx = F.Context(0)
blobs = synth_blobs(n)
roots, last = ctx.fri_commit_batch(blobs, None, F.PcsConfig(4, 0, 20, 20))
roots, last = ctx.fri_commit_batch(blobs, None, F.PcsConfig(4, 0, 20, 20))
print(roots[0, 0].tobytes().hex())
