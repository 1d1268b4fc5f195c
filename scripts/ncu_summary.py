#!/usr/bin/env python
"""Summarise an `ncu --set full` capture into a small CSV (the committed evidence under profiles/).
usage: python scripts/ncu_summary.py <x.ncu-rep | x_raw.csv (ncu -i x.ncu-rep --page raw --csv)> profiles/x_summary.csv"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in KEYS if k in idx]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keys)
        w.writerow([units[idx[k]] for k in keys])
        for r in data:
            w.writerow([r[idx[k]] for k in keys])
    print("wrote", out, len(data), "kernels")


if __name__ == "__main__":
    main()
