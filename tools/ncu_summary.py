#!/usr/bin/env python
"""Summarise an .ncu-rep (read here on the CPU box): per captured launch the metrics the roofline needs.

    python tools/ncu_summary.py gpurun_out/prof_full.ncu-rep > profiles/rNN_ncu_summary.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem"),
    ("launch__occupancy_limit_registers", "occ_lim_regs"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
    ("lts__t_sectors_op_atom.sum", "l2_atom_sectors"),
    ("lts__t_sectors_op_red.sum", "l2_red_sectors"),
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("| # | kernel | grid | block | " + " | ".join(n for _, n in WANT) + " |")
    print("|---|---|---|---|" + "---|" * len(WANT))
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0]
        cells = []
        for m, _ in WANT:
            if m in col:
                v, u = r[col[m]], units[col[m]]
                try:
                    f = float(v.replace(",", ""))
                    v = f"{f:,.0f}" if abs(f) >= 1000 else f"{f:.2f}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            else:
                cells.append("n/a")
        print(f"| {r[col['ID']]} | {name} | {r[col['Grid Size']]} | {r[col['Block Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
