#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the per-kernel lines kept under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "msecond": 1e-3, "nsecond": 1e-9}


def main(rep, hbm_gbs=6453.1):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[col["Kernel Name"]]
        print(f"kernel: {name[:110]}")
        vals = {}
        for k in KEYS:
            if k in col:
                v, u = r[col[k]], units[col[k]]
                print(f"  {k:<82s} {v:>16s} {u}")
                try:
                    vals[k] = float(v.replace(",", "")) * UNIT.get(u, 1.0)
                except ValueError:
                    pass
        if "dram__bytes_read.sum" in vals and "gpu__time_duration.sum" in vals:
            tot = vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]
            bw = tot / vals["gpu__time_duration.sum"] / 1e9
            print(f"  => DRAM traffic per launch {tot / 1e6:.1f} MB; {bw:.0f} GB/s = {bw / hbm_gbs:.2f} of the measured copy bandwidth")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
