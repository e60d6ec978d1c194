"""Summarise an ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) into the handful of numbers the
roofline discussion in DESIGN.md uses.  Usage: python profiles/summarize.py profiles/<file>_raw.csv"""
import csv
import sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    for v in vals:
        name = v[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
        print("kernel:", name[:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:70s} {v[i]} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and v[i]:
                stalls.append((float(v[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace(
                    "_per_issue_active.ratio", "")))
        print("  stalls per issue:", ", ".join(f"{n}={x:.2f}" for x, n in sorted(stalls, reverse=True)[:7]))


if __name__ == "__main__":
    main(sys.argv[1])
