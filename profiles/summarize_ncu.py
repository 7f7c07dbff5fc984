#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small per-kernel CSV kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep profiles/<name>_summary.csv
"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(w, hdr.index(w)) for w in WANT if w in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{n} [{units[i]}]" if units[i] else n for n, i in cols])
        for r in rows[2:]:
            w.writerow([r[i] for _, i in cols])
    print("wrote", out, len(rows) - 2, "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
