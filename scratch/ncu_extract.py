#!/usr/bin/env python
"""Turn an ncu report into the per-kernel summary kept under profiles/.

    python scratch/ncu_extract.py gpurun_out/x.ncu-rep profiles/rNN_ncu_full_x.csv

Reads `ncu -i REP --page raw --csv` (no GPU needed) and writes, per profiled launch: time, DRAM bytes
read/written and the achieved DRAM rate against MEASURED_PEAKS.json, registers, occupancy limiter,
issue/L1/L2 utilisation, warp instructions and the stall reasons per issued instruction."""
import csv
import json
import os
import subprocess
import sys

WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    peak = None
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        peak = float(json.load(open(os.path.join(here, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([k for k, _ in idx] + ["dram_GBps", "dram_frac_of_measured_peak"])
        w.writerow([units[i] for _, i in idx] + ["GB/s", ""])
        for d in data:
            col = {k: (d[i], units[i]) for k, i in idx}
            t_us = float(col["gpu__time_duration.sum"][0]) * TO_US.get(col["gpu__time_duration.sum"][1], 1.0)
            byts = sum(float(col[k][0]) * TO_BYTES.get(col[k][1], 1) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in col)
            gbps = byts / (t_us * 1e-6) / 1e9 if t_us else 0.0
            w.writerow([d[i] for _, i in idx] + ["%.1f" % gbps, "%.3f" % (gbps / peak) if peak else ""])
            print("%-70s %9.1f us %8.1f MB  %7.1f GB/s%s" % (col["Kernel Name"][0][:70], t_us, byts / 1e6, gbps,
                                                          "  (%.2f of peak)" % (gbps / peak) if peak else ""))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
