"""Builds and runs the re-expressed criterion benches (benches/commit.cpp, benches/proof.cpp) on cuda:0 and,
beside them, the CPU oracle port on the same inputs (one thread, like the reference), as a table.
Usage: python benches/run.py [--seconds S] [--no-cpu]      (needs a GPU; run under gpurun)"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(name):
    from frieda_b200 import build as fb
    fb.build()
    libdir = os.path.dirname(fb.LIB)
    out = os.path.join(ROOT, "benches", name)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "benches", name + ".cpp"), "-o", out, "-L", libdir, "-lfrieda_b200",
                           f"-Wl,-rpath,{libdir}"])
    return out


def cpu_rows(seconds):
    """The oracle port on the same inputs: test infrastructure timed as the CPU baseline (kind "port")."""
    from oracle import oracle as O
    cfg = O.make_config(4, 0, 20, 20)
    datas = [bytes(i % 256 for i in range(n)) for n in (1024, 4096, 16384, 65536)]
    with open(os.path.join(ROOT, "tests", "golden", "blob"), "rb") as f:
        datas.append(f.read())
    rows = {}

    def timeit(fn):
        fn()
        t, n = 0.0, 0
        while t < seconds or n < 2:
            t0 = time.perf_counter()
            fn()
            t += time.perf_counter() - t0
            n += 1
        return t / n * 1e6

    for d in datas:
        rows[("commit", len(d))] = timeit(lambda: O.commit(d, 4))
        rows[("commit_and_generate_proof", len(d))] = timeit(lambda: O.prove(d, len(d), cfg))
        rows[("generate_proof", len(d))] = rows[("commit_and_generate_proof", len(d))]
        _, pr = O.prove(d, len(d), cfg)
        rows[("verify_proof", len(d))] = timeit(lambda: O.verify(pr, len(d)))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=1.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    env = dict(os.environ, FRIEDA_BENCH_SECONDS=str(args.seconds))
    blob = os.path.join(ROOT, "tests", "golden", "blob")
    rows = []
    for name in ("commit", "proof"):
        exe = build(name)
        out = subprocess.run([exe, blob], env=env, capture_output=True, text=True, check=True).stdout
        rows += [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    cpu = {} if args.no_cpu else cpu_rows(min(args.seconds, 1.0))
    print(f"{'group':28s} {'bytes':>8s} {'GPU mean us':>12s} {'GPU min us':>11s} {'CPU port us':>12s} {'ratio':>8s}")
    for r in rows:
        c = cpu.get((r["group"], r["param"]))
        print(f"{r['group']:28s} {r['param']:8d} {r['mean_us']:12.1f} {r['min_us']:11.1f} "
              f"{(f'{c:12.1f}' if c else ' ' * 12)} {(f'{c / r_mean(r):8.1f}' if c else '')}")
    print(json.dumps({"criterion": rows, "cpu_port_us": {f"{k[0]}/{k[1]}": v for k, v in cpu.items()}}))


def r_mean(r):
    return r["mean_us"]


if __name__ == "__main__":
    main()
