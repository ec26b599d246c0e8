"""Summarise an .ncu-rep (ncu --set full) into a small CSV for profiles/: one row per captured launch.

    python tools/ncu_summary.py gpurun_out/r01c_prof.ncu-rep > profiles/r01c_ncu_full_summary.csv
"""
import csv
import io
import subprocess
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__cycles_elapsed.avg.per_second", "lts__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEEP if k in hdr]
    out = csv.writer(sys.stdout)
    out.writerow([hdr[i] for i in idx])
    out.writerow([units[i] for i in idx])
    for r in data:
        out.writerow([r[i] for i in idx])


if __name__ == "__main__":
    main()
