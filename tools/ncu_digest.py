#!/usr/bin/env python
"""Prints the handful of metrics we read from an `ncu --set full` report:  python tools/ncu_digest.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        print("=" * 100)
        for k in WANT:
            if k in d:
                print("%-70s %s %s" % (k, d[k], units[hdr.index(k)]))
        stalls = sorted(((float(d[k].replace(",", "")), k[len(STALL):-len("_per_issue_active.ratio")]) for k in hdr
                         if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and d[k] not in ("", "n/a")), reverse=True)
        print("top stalls (warps per issue):", ", ".join("%s=%.2f" % (n, v) for v, n in stalls[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
