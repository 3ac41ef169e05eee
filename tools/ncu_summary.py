"""Summarise an .ncu-rep (read here on the CPU box) into a small CSV for profiles/."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([f"{k} [{units[i]}]" if units[i] else k for k, i in idx])
        for r in rows[2:]:
            w.writerow([r[i] for _, i in idx])
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
