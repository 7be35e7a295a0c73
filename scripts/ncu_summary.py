"""Summarises an ncu report (.ncu-rep) into the handful of metrics DESIGN.md / profiles/ quote.
Usage: python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...]"""
import csv
import subprocess
import sys

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
