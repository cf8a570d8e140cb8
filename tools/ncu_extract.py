#!/usr/bin/env python
"""Turn `.ncu-rep` captures (gpurun_out/, scratch) into small tracked CSV extracts under profiles/.

  python tools/ncu_extract.py gpurun_out/prof_scan3d.ncu-rep profiles/ncu_r02_scan3d.csv

Keeps, per profiled launch, the metrics every number quoted in profiles/*.md and DESIGN.md comes from: duration,
DRAM bytes and throughput, FP64 pipe utilisation, issue-slot utilisation, warps active, registers, launch geometry.
Runs `ncu -i <rep> --page raw --csv` (the recipe of /opt/skills/guides/B200_PROFILING.md) and filters its columns.
"""
import csv
import io
import subprocess
import sys

KEEP = (
    "Kernel Name", "Block Size", "Grid Size",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum",
    "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # header row = the first one that names "Kernel Name"; the row after it holds the units
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[h], rows[h + 1]
    cols = [i for i, n in enumerate(names) if n in KEEP]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["source: ncu -i %s --page raw --csv (filtered by tools/ncu_extract.py)" % rep])
        w.writerow([names[i] for i in cols])
        w.writerow([units[i] for i in cols])
        for r in rows[h + 2:]:
            if len(r) == len(names):
                w.writerow([r[i] for i in cols])
    print("%s: %d launches, %d metrics -> %s" % (rep, len(rows) - h - 2, len(cols), out))


if __name__ == "__main__":
    main()
