"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [min_share_pct]"""
import csv
import subprocess
import sys


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                         capture_output=True, text=True).stdout
    cur, rows = None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] not in ("Function Name", "Line No", "") and r[0].isdigit():
            try:
                rows.append((cur, int(r[0]), r[1].strip(), int(r[7]), int(r[6])))
            except ValueError:
                pass
    tot = sum(x[3] for x in rows) or 1
    stot = sum(x[4] for x in rows) or 1
    print(f"total warp instructions {tot}, samples {stot}")
    for f, ln, src, n, smp in sorted(rows, key=lambda x: -x[3]):
        if 100.0 * n / tot >= thr:
            print(f"{100.0 * n / tot:5.1f}% inst {100.0 * smp / stot:5.1f}% smp  {f}:{ln:<4d} {src[:110]}")


if __name__ == "__main__":
    main()
