"""Summarise an .ncu-rep (raw page) into one row per kernel launch. Usage: python tools/ncu_summary.py report.ncu-rep [out.csv]"""
import csv
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "warp_inst"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit"), ("lts__t_sector_hit_rate.pct", "l2_hit")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    w = csv.writer(open(sys.argv[2], "w", newline="") if len(sys.argv) > 2 else sys.stdout)
    w.writerow(["id", "kernel", "grid", "block"] + [f"{n} [{units[hdr.index(c)]}]" for c, n in COLS])
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].replace("void ", "").replace("psb::", "").replace("<unnamed>::", "").split("(")[0]
        w.writerow([r[hdr.index("ID")], name, r[hdr.index("Grid Size")], r[hdr.index("Block Size")]] + [r[hdr.index(c)] for c, _ in COLS])


if __name__ == "__main__":
    main()
