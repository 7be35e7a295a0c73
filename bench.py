#!/usr/bin/env python
"""bench.py -- FRIEDA commit-path throughput on B200 (BASELINE.json metric: blobs committed/s and
input GB/s vs the reference CPU path; % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = one pass of the hot path (pack -> circle-FFT LDE -> Merkle -> FRI layers with their
commitments, i.e. FriProver::commit of src/proof.rs:38-57) over one batch of synthetic blobs:
BASELINE config 3, `--blobs` independent 128 KiB blobs PER GPU (weak scaling, no collectives).

  value : whole-job blobs/s with the inputs already resident in HBM (frieda_fri_commit_batch_device)
  e2e   : same metric through the host-buffer C-ABI call (frieda_fri_commit_batch): pinned host
          input -> H2D -> kernels -> D2H of the layer roots and last-layer polynomials, all timed
  roofline     : the dominant kernel (Merkle bottom pass) against the measured HBM peak
  passes       : the LDE and fold passes timed standalone against the HBM peak (BASELINE target)
  cpu_baseline : the CPU oracle (a port of the reference's CpuBackend path) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOB_LEN = 131072                     # EIP-4844 blob size (BASELINE configs 2-4)
CFG = (4, 0, 20, 20)                  # benches/proof.rs:5-12: blowup 2^4, log_last 0, 20 queries, pow 20
SPLITMIX0 = 0x4652494544410000
METRIC = "blobs_committed_per_s"
UNIT = "blobs/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # under load = samples in the upper half of the observed power range
        thr = (min(power) + max(power)) / 2 if power else 0
        loaded = [s for s, p in zip(sm, power) if p >= thr] or sm
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def synth_blobs(n, rank_offset=0):
    """n blobs of SplitMix64 bytes, blob b seeded with SPLITMIX0 + b -> (n, BLOB_LEN) uint8."""
    import numpy as np
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    words = BLOB_LEN // 8
    with np.errstate(over="ignore"):
        idx = np.arange(1, words + 1, dtype=np.uint64)[None, :]
        s0 = (np.uint64(SPLITMIX0) + np.arange(rank_offset, rank_offset + n, dtype=np.uint64))[:, None]
        z = (s0 + idx * np.uint64(0x9E3779B97F4A7C15)) & M
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.astype("<u8").view(np.uint8).reshape(n, BLOB_LEN)


# ------------------------------------------------------------------------------------------
def cpu_sample(n_threads: int, blobs_per_thread: int, results=None):
    """Times the CPU oracle (FRI commit phase per blob) on n_threads host threads.  Thread i works on
    synthetic blob i; `results[i]` receives its layer roots (the checker side of the parity spot check)."""
    import numpy as np
    from oracle import oracle as O
    O.lib()
    cfg = O.make_config(*CFG)
    data = [synth_blobs(1, i)[0].tobytes() for i in range(n_threads)]
    errs = []

    def work(i):
        try:
            for _ in range(blobs_per_thread):
                roots, _ = O.fri_commit(data[i], None, cfg)
            if results is not None:
                results[i] = roots
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ths = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    if errs:
        raise errs[0]
    return n_threads * blobs_per_thread / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores.  The Rust
    crate cannot be built here (no cargo, stwo un-vendored), so this is the oracle port, which
    mirrors stwo's CpuBackend structure (oracle/frieda_oracle.c header)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_threads = max(1, min(cores, 64))
    for _ in range(args.warmup):
        cpu_sample(n_threads, 1)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        v, dt = cpu_sample(n_threads, 1)
        t_total += dt
        n_total += n_threads
    value = n_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (M31) / u32 (BLAKE2s)",
        "data": "synthetic", "input_gb_per_s": value * BLOB_LEN / 1e9,
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": n_threads, "kind": "port",
                         "sample": f"{n_threads} blobs per step (one per thread), {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def compressions_per_blob(blob_len, cfg):
    """BLAKE2s compressions of commit + FRI layers for one blob, from the geometry: every committed layer
    of 2^d points is a full binary tree (2^(d+1) - 1 compressions); layers d = D .. log_last + log_blowup + 1."""
    n_felts = (blob_len * 8 + 29) // 30
    poly_log = max((n_felts - 1).bit_length(), 2) - 2
    D = poly_log + cfg[0]
    last_log = cfg[1] + cfg[0]
    return sum((2 << d) - 1 for d in range(last_log + 1, D + 1))


def ncu_metrics():
    """{kernel name as ctx profiling reports it: record} from the newest tracked profiles/r*_ncu_metrics.json."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_metrics.json")))
    if not files:
        return {}
    try:
        with open(files[-1]) as f:
            d = json.load(f)
    except Exception:
        return {}
    for rec in d.values():
        rec["file"] = os.path.relpath(files[-1], ROOT)
    return d


def workload_config(args, blobs_per_step=None):
    return {
        "workload": "C3: batch of independent 128 KiB blobs, commit + FRI layers "
                    "(FriProver::commit: LDE, Merkle, 13 inner layers, last-layer poly)",
        "blob_bytes": BLOB_LEN, "blobs_per_gpu_per_step": blobs_per_step if blobs_per_step else args.blobs,
        "log_blowup": CFG[0], "log_last_layer_degree_bound": CFG[1], "parallelism": f"blob-sharded x{args.gpus}, no collectives",
        "l2": "per-step input (512 MiB at 4096 blobs) and working set exceed the 126 MB L2",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="frieda_b200", choices=["frieda_b200", "reference"])
    ap.add_argument("--blobs", type=int, default=4096, help="blobs per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-passes", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import frieda_b200 as F

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    # stdout carries the one JSON line and nothing else: libraries that write to fd 1 (NCCL announces its
    # version there when its first communicator is created) go to stderr until the line is printed
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: frieda_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hbm_peak, peak_src = measured_peaks()

    ctx = F.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local))
    cfg = F.PcsConfig(*CFG)
    n = args.blobs
    L = 1 + ctx.n_inner_layers(BLOB_LEN, cfg)
    host_np = synth_blobs(n, rank * n)
    h_in = torch.from_numpy(host_np).pin_memory()
    d_in = h_in.cuda(non_blocking=False)
    d_roots = torch.zeros((n, L, 32), dtype=torch.uint8, device="cuda")
    d_last = torch.zeros((n, 1, 4), dtype=torch.int32, device="cuda")
    h_roots = torch.zeros((n, L, 32), dtype=torch.uint8).pin_memory()
    h_last = torch.zeros((n, 1, 4), dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    def step_device():
        ctx.fri_commit_batch_ptr(d_in.data_ptr(), BLOB_LEN, BLOB_LEN, n, None, cfg, d_roots.data_ptr(),
                                 d_last.data_ptr(), device=True)

    def step_host():
        ctx.fri_commit_batch_ptr(h_in.data_ptr(), BLOB_LEN, BLOB_LEN, n, None, cfg, h_roots.data_ptr(),
                                 h_last.data_ptr(), device=False)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if profile:
            ctx.profile_read(reset=True)
            ctx.set_profiling(True)
        launches0 = ctx.launch_count
        ev0.record(stream)
        for _ in range(steps):
            fn()
        ev1.record(stream)
        stream.synchronize()
        barrier()
        ms = ev0.elapsed_time(ev1)
        prof = None
        if profile:
            ctx.set_profiling(False)
            prof = ctx.profile_read(reset=True)
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), ctx.launch_count - launches0, prof

    # parity spot check before timing: blob 0 of this rank against the golden/oracle-free invariant
    # (first-layer root == commit root through the other entry point)
    step_device()
    stream.synchronize()
    root0 = bytes(d_roots[0, 0].cpu().numpy())
    assert root0 == ctx.commit(host_np[0], CFG[0]), "fri_commit layer-0 root != commit root"

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, launches, prof = timed(step_device, args.steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_host()
    ms_e2e, _, _ = timed(step_host, args.steps)
    assert bytes(h_roots[0, 0].numpy()) == root0, "host-path root differs from device-path root"

    total_blobs = n * world * args.steps
    value = total_blobs / (ms_dev / 1e3)
    e2e_value = total_blobs / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel, from the event-timed launches of the timed region
    N = 1 << 18
    dominant = max(prof.items(), key=lambda kv: kv[1][1]) if prof else (None, (0, 0.0))
    dom_name, (dom_launches, dom_ms) = dominant
    # algorithmic bytes per launch: the depth-3 bottom pass over layer 0 reads the 4
    # evaluation columns (16 B per leaf) and writes one 32-B node per 8 leaves, per blob
    alg_bytes = {"merkle_bottom_cols": n * (16 * N + 32 * (N >> 3)),
                 "fold_circle+merkle_bottom": n * (16 * N + 8 * N + 32 * (N >> 4)),
                 }.get(dom_name)
    # DRAM traffic of the same kernel: dram__bytes_read.sum + dram__bytes_write.sum of the tracked `ncu --set full`
    # capture (profiles/r*_ncu_metrics.json, written by scripts/ncu_summary.py --json with the git hash and the blobs
    # per launch of the capture), scaled per blob
    ncu_rec = ncu_metrics().get(dom_name)
    ncu_traffic_per_blob = ((ncu_rec["dram_bytes_read"] + ncu_rec["dram_bytes_write"]) / ncu_rec["blobs_per_launch"]
                            if ncu_rec else None)
    roofline = None
    if dom_name and dom_launches and alg_bytes:
        per_launch_s = dom_ms / 1e3 / dom_launches
        achieved = alg_bytes / per_launch_s / 1e9
        roofline = {"kernel": dom_name, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                    "frac": achieved / hbm_peak,
                    "traffic": ncu_traffic_per_blob * n if ncu_traffic_per_blob else None,
                    "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                    "avg_launch_ms": dom_ms / dom_launches, "share_of_step": dom_ms / ms_dev,
                    "traffic_source": ({k: ncu_rec[k] for k in ("file", "git", "blobs_per_launch")} if ncu_rec else None),
                    "note": "this kernel is INT ALU-pipe bound (BLAKE2s), not HBM bound: see int_roofline; the HBM "
                            "roof applies to the LDE / fold passes under `passes`. traffic = ncu DRAM bytes per blob "
                            "of the tracked capture scaled to this launch's blob count"}
    kernels = {k: {"launches": v[0], "total_ms": round(v[1], 3), "share": round(v[1] / ms_dev, 4)}
               for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])} if prof else {}
    # integer roofline: compressions per blob (SURVEY 8d, C2 commit+FRI = 1,048,498).  Each needs 648
    # xor/rotate instructions that only the ALU pipe executes, at 0.5 warp-instr/clk/SMSP (measured,
    # profiles/r01_pipe_rates_b200.txt): peak = 592 SMSPs x 32 lanes x clk / (648 x 2)
    compress_per_blob = compressions_per_blob(BLOB_LEN, CFG)
    assert compress_per_blob == 1048498  # SURVEY 8(d), C2 commit + FRI
    hashes_per_s = value / world * compress_per_blob
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    int_peak = 148 * 4 * 32 * sm_mhz * 1e6 / (648 * 2)
    int_roofline = {"bound": "int_alu_pipe", "achieved": hashes_per_s, "peak": int_peak, "unit": "compressions/s",
                    "frac": hashes_per_s / int_peak, "alu_instr_per_compression_floor": 648, "sm_mhz_used": sm_mhz,
                    "note": "whole step incl. LDE/fold time; the hash kernels alone run at ~0.90 of this roof (columns pass 0.92)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 (M31 field) / u32 (BLAKE2s)", "data": "synthetic",
        "input_gb_per_s": value * BLOB_LEN / 1e9,
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * BLOB_LEN,
                "d2h_bytes_per_step": n * (L * 32 + 16), "ms_per_step": ms_e2e / args.steps,
                "input_gb_per_s": e2e_value * BLOB_LEN / 1e9},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "int_roofline": int_roofline,
        "kernels": kernels,
    }

    # free the big C3 buffers before the extra measurements
    del d_in, d_roots, d_last
    torch.cuda.empty_cache()
    c5 = None
    if not args.no_extras:
        # BASELINE config 5: ONE 64 MiB blob at blowup 2^2 split over all ranks (subtree roots exchanged over peer
        # memory); strong scaling.  Every rank takes part; timed end to end (H2D of the blob included); the root is
        # ASSERTED against the oracle's on every rank.
        c5 = c5_split(ctx, stream, torch, dist, distributed, rank, world, barrier)
        line["c5_split"] = c5
    if rank == 0 and not args.no_extras:
        line["extras"] = extras(ctx, h_in.numpy(), cfg, torch)  # pinned host memory
    if rank == 0 and not args.no_passes:
        line["passes"] = standalone_passes(ctx, stream, torch, np, hbm_peak, peak_src)
        line["single_blob_latency_ms"] = single_blob_latency(ctx, host_np, cfg)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:  # reported at N = 1 only
        cores = os.cpu_count() or 1
        n_threads = max(1, min(cores, 64))
        v1, t1 = cpu_sample(1, 1)
        per_thread = max(1, min(8, int(round(12.0 / max(t1, 1e-3)))))
        oracle_roots = {}
        v, dt = cpu_sample(n_threads, per_thread, oracle_roots)
        # the baseline run doubles as a parity check: the oracle's layer roots of blobs 0..n_threads-1
        # against what the timed GPU path produced for the same blobs
        gpu_roots = h_roots.numpy()
        same = all(b"".join(oracle_roots[i]) == gpu_roots[i].tobytes() for i in oracle_roots)
        line["parity_vs_oracle"] = {"blobs_checked": len(oracle_roots), "layer_roots_per_blob": L, "all_equal": same}
        assert same, "GPU layer roots differ from the oracle"
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": n_threads, "kind": "port",
                                "sample": f"{n_threads * per_thread} blobs ({per_thread} per thread) in {dt:.1f} s; "
                                          f"1 thread: {v1:.3f} blobs/s",
                                "single_thread_value": v1}
    # compact copies of the two other BASELINE configs as the LAST keys, so that a tail of the line carries them
    if rank == 0 and "extras" in line:
        pc = line["extras"]["prove_c4_e2e"]
        line["prove_c4"] = {"proofs_per_s": round(pc["blobs_per_s"], 1), "blobs": pc["blobs"], "n_queries": 64,
                            "verify_ok": pc["proofs_verify"]}
    if c5 is not None and "ms_per_blob" in c5:
        line["c5"] = {"ms": round(c5["ms_per_blob"], 4), "n": world, "root_ok": c5["root_matches_oracle"],
                      "ms_resident": (round(c5["ms_per_blob_slices_resident"], 4)
                                      if c5.get("ms_per_blob_slices_resident") else None),
                      "exchange": c5["exchange_short"], "gb_per_s": round(c5["input_gb_per_s"], 2),
                      "fri_ms": round(c5["fri_commit_split_ms"], 3) if c5.get("fri_commit_split_ms") else None,
                      "prove_ms": round(c5["prove_split_ms"], 3) if c5.get("prove_split_ms") else None,
                      "split_ok": c5.get("split_fri_and_proof_ok")}
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
    return 0


C5_ROOT = "7ab1d35e23dcb4b678524912e0fc0cdbcb6efaa34aa9f480f97f3138b3f0ee07"  # oracle, tests/golden/vectors.json


def c5_split(ctx, stream, torch, dist, distributed, rank, world, barrier):
    import numpy as np
    from frieda_b200.parallel import commit_split, is_pow2
    if not is_pow2(world):
        return {"skipped": "world size is not a power of two"}
    n_bytes = 64 << 20
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        idx = np.arange(1, n_bytes // 8 + 1, dtype=np.uint64)
        z = (np.uint64(0x4652494544414236) + idx * np.uint64(0x9E3779B97F4A7C15)) & M
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    blob = torch.from_numpy(z.astype("<u8").view(np.uint8)).pin_memory().numpy()
    root = commit_split(ctx, blob, 2, rank=rank, world=world)
    ok = root.hex() == C5_ROOT
    assert ok, f"rank {rank}/{world}: split commit root {root.hex()} != oracle root {C5_ROOT}"
    for _ in range(3):
        commit_split(ctx, blob, 2, rank=rank, world=world)
    iters = 8
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(iters):
        ok = ok and commit_split(ctx, blob, 2, rank=rank, world=world).hex() == C5_ROOT
    ev1.record(stream)
    stream.synchronize()
    wall = time.perf_counter() - t0
    assert ok, f"rank {rank}/{world}: a timed split commit returned a different root"
    dt = torch.tensor([ev0.elapsed_time(ev1) / 1e3, wall], dtype=torch.float64, device="cuda")
    if distributed:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    per = float(dt[0].item()) / iters
    wall_per = float(dt[1].item()) / iters
    # the same with every rank's slice already in its symmetric buffer (inputs resident in HBM, like `value`)
    resident_per = None
    from frieda_b200 import parallel as _par
    if world > 1 and not _par._peer_memory_broken:
        from frieda_b200.parallel import commit_split_peers
        host = blob.reshape(-1)
        assert commit_split_peers(ctx, host, host.size, 2, rank, world, resident=True).hex() == C5_ROOT
        barrier()
        ev0.record(stream)
        for _ in range(iters):
            ok = ok and commit_split_peers(ctx, host, host.size, 2, rank, world, resident=True).hex() == C5_ROOT
        ev1.record(stream)
        stream.synchronize()
        assert ok, f"rank {rank}/{world}: a split commit from resident slices returned a different root"
        dr = torch.tensor([ev0.elapsed_time(ev1) / 1e3], dtype=torch.float64, device="cuda")
        dist.all_reduce(dr, op=dist.ReduceOp.MAX)
        resident_per = float(dr.item()) / iters
    # the FRI commit phase and a whole proof of the same blob, every layer split over the ranks (frieda_fri_split_*).
    # A widening row: its checks are REPORTED (split_fri_and_proof_ok), every rank agreeing on the
    # outcome first, so that a failure here cannot take the C3 / C5 numbers of the line down with it.
    from frieda_b200.parallel import fri_commit_split, prove_split
    import frieda_b200 as F
    c5cfg = F.PcsConfig(2, 0, 64, 20)
    fri_ms = prove_ms = None
    split_ok, split_err = False, None
    try:
        roots, _ = fri_commit_split(ctx, blob, None, c5cfg)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            roots2, _ = fri_commit_split(ctx, blob, None, c5cfg)
        torch.cuda.synchronize()
        df = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        proot, proof = prove_split(ctx, blob, 1, c5cfg)
        barrier()
        t0 = time.perf_counter()
        proot, proof = prove_split(ctx, blob, 1, c5cfg)
        torch.cuda.synchronize()
        dp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if distributed:
            dist.all_reduce(df, op=dist.ReduceOp.MAX)
            dist.all_reduce(dp, op=dist.ReduceOp.MAX)
        fri_ms, prove_ms = float(df.item()) / 3 * 1e3, float(dp.item()) * 1e3
        split_ok = bool(roots[0].tobytes().hex() == C5_ROOT and np.array_equal(roots, roots2) and
                        proot.hex() == C5_ROOT and F.verify_proof(proof, 1) and not F.verify_proof(proof, 2))
    except Exception as e:  # noqa: BLE001 -- reported in the line
        split_err = f"{type(e).__name__}: {e}"[:200]
    if distributed:
        flag = torch.tensor([1 if split_ok else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        split_ok = bool(flag.item())
    exchange = ("single GPU" if world == 1 else
                "NCCL all-gathers (peer mapping unavailable)" if _par._peer_memory_broken else
                "peer-mapped memory: slices and roots read in place over NVLink by the library's kernels")
    return {"workload": "C5: one 64 MiB blob, blowup 2^2, commit split into per-GPU subtrees, subtree roots combined",
            "exchange": exchange,
            "exchange_short": ("single" if world == 1 else "nccl" if _par._peer_memory_broken else "peer"),
            "root_matches_oracle": ok, "ms_per_blob": per * 1e3, "wall_ms_per_blob": wall_per * 1e3,
            "ms_per_blob_slices_resident": resident_per * 1e3 if resident_per else None,
            "fri_commit_split_ms": fri_ms, "prove_split_ms": prove_ms, "split_fri_and_proof_ok": split_ok,
            "split_fri_error": split_err,
            "split_fri_note": "FriProver::commit / commit_and_generate_proof of the same blob (blowup 2^2, 64 queries, "
                              "pow 20) with every layer split over the ranks; host wall clock, max over ranks, H2D of the "
                              "blob included; layer-0 root == the oracle's commit root, proof verified by the host verifier",
            "blobs_per_s": 1.0 / per,
            "input_gb_per_s": n_bytes / per / 1e9, "n_gpus": world, "scaling": "strong",
            "timing": "CUDA events on the context's stream around the synchronous calls, max over ranks (H2D of the "
                      "blob included); wall_ms_per_blob = host clock around the same region"}


def extras(ctx, host_np, cfg, torch):
    """Other entry points on the same synthetic blobs (rank 0): commit only (api::commit, batched) and
    full proof generation with 64 queries per blob (BASELINE config 4)."""
    import frieda_b200 as F
    out = {}
    n = min(len(host_np), 4096)
    blobs = host_np[:n]
    ctx.commit_batch(blobs, CFG[0])
    t0 = time.perf_counter()
    ctx.commit_batch(blobs, CFG[0])
    dt = time.perf_counter() - t0
    out["commit_only_e2e"] = {"blobs": n, "blobs_per_s": n / dt, "input_gb_per_s": n * BLOB_LEN / dt / 1e9}
    npv = min(n, 512)
    c4 = F.PcsConfig(CFG[0], CFG[1], 64, CFG[3])
    seeds = list(range(npv))
    ctx.prove_batch(blobs[:npv], seeds, c4)  # warm-up at the same size (workspace allocation)
    # the per-kernel breakdown comes from a separate, instrumented call (two CUDA events around every launch); the
    # rate is timed on plain calls (median of three wall-clock measurements of the whole synchronous call)
    ctx.profile_read(reset=True)
    ctx.set_profiling(True)
    ctx.prove_batch(blobs[:npv], seeds, c4)
    ctx.set_profiling(False)
    prof = ctx.profile_read(reset=True)
    dts = []
    roots = proofs = None
    for _ in range(3):
        del roots, proofs  # a caller consumes its proofs before asking for more: their memory is free again
        t0 = time.perf_counter()
        roots, proofs = ctx.prove_batch(blobs[:npv], seeds, c4)
        dts.append(time.perf_counter() - t0)
    dt = sorted(dts)[1]
    ok = all(F.verify_proof(proofs[i], seeds[i]) for i in (0, npv // 2, npv - 1))
    # GPU batch verification of those proofs next to the host verifier
    many = (proofs * ((4096 + npv - 1) // npv))[:4096]
    many_seeds = (seeds * ((4096 + npv - 1) // npv))[:4096]
    ctx.verify_batch(many[:256], many_seeds[:256])
    t0 = time.perf_counter()
    res = ctx.verify_batch(many, many_seeds)
    dtv = time.perf_counter() - t0
    t0 = time.perf_counter()
    host_ok = all(F.verify_proof(proofs[i], seeds[i]) for i in range(32))
    dth = (time.perf_counter() - t0) / 32
    # same proofs as serialised bytes (what a light client holds): kernels timed with CUDA events
    import numpy as np
    pieces = [p.serialize() for p in proofs]
    offs1 = np.concatenate([[0], np.cumsum([len(b) for b in pieces])]).astype(np.uint64)
    reps = (4096 + npv - 1) // npv
    blob = torch.from_numpy(np.tile(np.frombuffer(b"".join(pieces), dtype=np.uint8), reps)).pin_memory().numpy()
    offs = np.concatenate([offs1[:-1] + r * offs1[-1] for r in range(reps)] + [[reps * offs1[-1]]]).astype(np.uint64)
    sb = (seeds * reps)
    ctx.verify_batch_bytes(blob, offs, sb)
    ctx.profile_read(reset=True)
    ctx.set_profiling(True)
    t0 = time.perf_counter()
    resb = ctx.verify_batch_bytes(blob, offs, sb)
    dtb = time.perf_counter() - t0
    ctx.set_profiling(False)
    vprof = ctx.profile_read(reset=True)
    kern_ms = sum(v[1] for v in vprof.values())
    out["verify_batch"] = {"proofs": len(many), "all_valid": all(r == 1 for r in res) and all(r == 1 for r in resb)
                           and host_ok,
                           "proofs_per_s_from_structs": len(many) / dtv, "proofs_per_s_from_bytes": len(resb) / dtb,
                           "kernel_ms": {k: round(v[1], 3) for k, v in vprof.items()},
                           "proofs_per_s_kernels_only": len(resb) / (kern_ms / 1e3) if kern_ms else None,
                           "proof_bytes_total": int(offs[-1]),
                           "host_verifier_proofs_per_s_1_thread": 1.0 / dth,
                           "note": "wall clock incl. upload of the proofs (132 KB each); from_structs also serialises"}
    out["prove_c4_e2e"] = {"blobs": npv, "n_queries": 64, "pow_bits": CFG[3], "blobs_per_s": npv / dt,
                           "proofs_verify": ok, "wall_ms": dt * 1e3,
                           "kernel_ms": {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                           "note": "commit + FRI + grind + decommit + host proof assembly, kept trees; wall = median of 3 plain "
                                   "calls, kernel_ms from a separate instrumented call"}
    return out


def standalone_passes(ctx, stream, torch, np, hbm_peak, peak_src):
    """LDE and fold passes timed alone (BASELINE target: >= 60% of the HBM roofline).  Working sets
    (>= 1 GiB) exceed L2, so no flush is needed between iterations."""
    out = {}
    nb, p, beta = 256, 14, 4
    n_felts = 34953
    D = p + beta
    rng = np.random.default_rng(0)
    coef = np.zeros((nb, 4 << p), dtype=np.uint32)
    coef[:, :n_felts] = rng.integers(0, (1 << 31) - 1, (nb, n_felts), dtype=np.uint32)
    d_coef = torch.from_numpy(coef.view(np.int32)).cuda()
    d_eval = torch.empty((nb, 4 << D), dtype=torch.int32, device="cuda")
    d_line = torch.empty((nb, 4 << (D - 1)), dtype=torch.int32, device="cuda")
    d_line2 = torch.empty((nb, 4 << (D - 2)), dtype=torch.int32, device="cuda")
    d_alpha = torch.from_numpy(rng.integers(0, (1 << 31) - 1, (nb, 4), dtype=np.uint32).view(np.int32)).cuda()
    torch.cuda.synchronize()

    def time_it(fn, iters=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream.synchronize()
        e0.record(stream)
        for _ in range(iters):
            fn()
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / iters

    N = 1 << D
    ms = time_it(lambda: ctx.pass_lde(d_coef.data_ptr(), p, beta, nb, n_felts, d_eval.data_ptr()))
    b = nb * (16 * (1 << p) + 16 * N)
    out["lde"] = {"algorithmic_bytes": b, "ms": ms, "achieved_gbs": b / ms / 1e6, "frac": b / ms / 1e6 / hbm_peak}
    ms = time_it(lambda: ctx.pass_fold(d_eval.data_ptr(), D, True, nb, d_alpha.data_ptr(), d_line.data_ptr()))
    b = nb * (16 * N + 8 * N)
    out["fold_circle"] = {"algorithmic_bytes": b, "ms": ms, "achieved_gbs": b / ms / 1e6,
                          "frac": b / ms / 1e6 / hbm_peak}
    ms = time_it(lambda: ctx.pass_fold(d_line.data_ptr(), D - 1, False, nb, d_alpha.data_ptr(), d_line2.data_ptr()))
    b = nb * (8 * N + 4 * N)
    out["fold_line"] = {"algorithmic_bytes": b, "ms": ms, "achieved_gbs": b / ms / 1e6,
                        "frac": b / ms / 1e6 / hbm_peak}
    out["peak_gbs"] = hbm_peak
    out["peak_source"] = peak_src
    out["shape"] = f"{nb} blobs x 4 columns, poly_log {p}, blowup 2^{beta} (C2 shape)"
    return out


def single_blob_latency(ctx, host_np, cfg):
    one = host_np[:1]
    for _ in range(3):
        ctx.fri_commit_batch(one, None, cfg)
    t0 = time.perf_counter()
    iters = 20
    for _ in range(iters):
        ctx.fri_commit_batch(one, None, cfg)
    return (time.perf_counter() - t0) / iters * 1e3


if __name__ == "__main__":
    sys.exit(main())
