"""Host-side stage trace + per-kernel times of single proofs at the criterion bench sizes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frieda_b200 as F
ctx = F.Context(0)
cfg = F.PcsConfig(4, 0, 20, 20)
for n in (1024, 65536, 262146):
    data = bytes(i % 256 for i in range(n))
    for _ in range(3):
        ctx.commit_and_generate_proof(data, n, cfg)
    ctx.set_profiling(True)
    t0 = time.perf_counter()
    _, pr = ctx.commit_and_generate_proof(data, n, cfg)
    dt = time.perf_counter() - t0
    ctx.set_profiling(False)
    prof = ctx.profile_read()
    print(f"n={n} wall {dt*1e6:.0f} us nonce {pr.proof_of_work} kernels {sum(v[1] for v in prof.values())*1e3:.0f} us launches {sum(v[0] for v in prof.values())}")
    print("  ", {k: (v[0], round(v[1] * 1e3, 1)) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])})
    os.environ["FRIEDA_TRACE"] = "1"
    ctx.commit_and_generate_proof(data, n, cfg)
    del os.environ["FRIEDA_TRACE"]
