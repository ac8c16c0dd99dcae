#!/usr/bin/env python
"""Summarise an ncu report (ncu -i X.ncu-rep --page raw --csv) into a small markdown table for profiles/."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def main(rep, out, pick="first"):
    if rep.endswith(".csv"):   # already exported with `ncu -i X.ncu-rep --page raw --csv`
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    seen = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]].split("(")[0].split("::")[-1]
        if pick == "longest":     # several shapes per kernel in one capture: keep the longest launch
            if name not in seen or float(r[ix["gpu__time_duration.sum"]].replace(",", "")) > float(seen[name][ix["gpu__time_duration.sum"]].replace(",", "")):
                seen[name] = r
        else:
            seen.setdefault(name, r)  # first captured launch of each kernel
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep}` ({pick} captured launch per kernel)\n\n")
        f.write("| metric | " + " | ".join(seen) + " |\n|---|" + "---|" * len(seen) + "\n")
        for k in KEYS:
            if k in ix:
                f.write(f"| {k} [{units[ix[k]]}] | " + " | ".join(seen[n][ix[k]] for n in seen) + " |\n")
    print(open(out).read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
