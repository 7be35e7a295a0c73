"""CUDA-event time of the standalone LDE pass with a checksum, at the bench's C2 shape (256 blobs, poly_log 14,
blowup 2^4), at C1's shape (poly_log 15) and at C5's shape (ONE column set of poly_log 23, blowup 2^2).
Environment switches read by the library are echoed so that A/B runs are self-describing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import frieda_b200 as F
ctx = F.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream_ptr)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FRIEDA_"))


def run(nb, p, beta, n_felts, iters):
    D = p + beta
    rng = np.random.default_rng(0)
    coef = np.zeros((nb, 4 << p), dtype=np.uint32)
    coef[:, :n_felts] = rng.integers(0, (1 << 31) - 1, (nb, n_felts), dtype=np.uint32)
    d_coef = torch.from_numpy(coef.view(np.int32)).cuda()
    d_eval = torch.empty((nb, 4 << D), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    f = lambda: ctx.pass_lde(d_coef.data_ptr(), p, beta, nb, n_felts, d_eval.data_ptr())
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream.synchronize()
    e0.record(stream)
    for _ in range(iters):
        f()
    e1.record(stream)
    stream.synchronize()
    ms = e0.elapsed_time(e1) / iters
    b = nb * (16 * (1 << p) + 16 * (1 << D))
    print(f"[{tag}] lde nb={nb} p={p} beta={beta}: {ms:.4f} ms  {b/ms/1e6:.0f} GB/s  {b/ms/1e6/6552:.3f} of 6552",
          "checksum", int(d_eval.view(torch.int64).sum().item()) & 0xffffffff)
    del d_coef, d_eval
    torch.cuda.empty_cache()


which = sys.argv[1] if len(sys.argv) > 1 else "all"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else None
if which in ("all", "c2"):
    run(256, 14, 4, 34953, iters or 20)
if which in ("all", "c1"):
    run(128, 15, 4, 69906, iters or 20)
if which in ("all", "c5"):
    run(1, 23, 2, 17895698, iters or 10)
