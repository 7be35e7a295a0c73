"""Summarises an ncu report (.ncu-rep) into the handful of metrics DESIGN.md / profiles/ quote.
Usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]
       python scripts/ncu_summary.py --json profiles/r02_ncu_metrics.json --blobs 296 [--git HASH] rep [rep ...]
The --json form writes the record bench.py reads for `roofline.traffic`: per kernel (named as the context's
profiler names it) the DRAM bytes of ONE launch, the blobs that launch processed and the git hash of the code."""
import csv
import json
import os
import subprocess
import sys

# ncu kernel name -> the name frieda_ctx_profile_read reports
PROFILE_NAMES = [
    ("merkle_bottom_kernel<0", "merkle_bottom_cols"), ("merkle_bottom_kernel<1", "fold_circle+merkle_bottom"),
    ("merkle_bottom_kernel<2", "fold_line+merkle_bottom"), ("merkle_bottom_kernel<3", "merkle_mid"),
    ("merkle_top_kernel", "merkle_top"), ("fri_tail_kernel", "fri_tail"), ("lde_warp_kernel", "lde"),
    ("lde_strided", "lde_strided"), ("fold_kernel", "fold"), ("pack_peers_kernel", "pack"), ("grind_kernel", "grind"),
]


def to_bytes(val, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(val.replace(",", "")) * scale


def write_json(out_path, blobs, git, reps):
    rec = {}
    if os.path.exists(out_path):
        with open(out_path) as f:
            rec = json.load(f)
    for path in reps:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            kname = vals[hdr.index("Kernel Name")]
            short = next((p for k, p in PROFILE_NAMES if k in kname), None)
            if short is None or short in rec and rec[short].get("report") != os.path.basename(path):
                continue
            ir, iw, it = (hdr.index(m) for m in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
            rec[short] = {"ncu_kernel": kname, "dram_bytes_read": to_bytes(vals[ir], units[ir]),
                          "dram_bytes_write": to_bytes(vals[iw], units[iw]), "time_ms_under_ncu": float(vals[it]) *
                          {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[it], 1),
                          "blobs_per_launch": blobs, "git": git, "report": os.path.basename(path)}
    with open(out_path, "w") as f:
        json.dump(rec, f, indent=1, sort_keys=True)
    print("wrote", out_path, sorted(rec))

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--json":
        args = sys.argv[2:]
        out_path = args.pop(0)
        blobs, git = 296, subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True,
                                         text=True).stdout.strip()
        while args and args[0].startswith("--"):
            flag = args.pop(0)
            if flag == "--blobs":
                blobs = int(args.pop(0))
            elif flag == "--git":
                git = args.pop(0)
        return write_json(out_path, blobs, git, args)
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            print(f"## {path}\nkernel: {name}")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    print(f"  {w:72s} {vals[i]:>18s} {units[i]}")
            stalls = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try:
                        stalls.append((float(vals[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("  top stalls (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]))
            print()


if __name__ == "__main__":
    main()
