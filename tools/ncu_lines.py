"""Per-source-line totals of one kernel from an .ncu-rep captured with --import-source on:
  ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:<name> > x.csv ; python tools/ncu_lines.py x.csv [min_pct]"""
import csv
import sys


def main(path, min_pct=0.7):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [r for r in rows if len(r) == len(hdr) and r[0].isdigit() and r[iS].isdigit()]
    tot = sum(int(r[iS]) for r in data)
    totE = sum(int(r[iE]) for r in data)
    print("samples", tot, "warp instructions executed", totE)
    for r in data:
        s, e = 100.0 * int(r[iS]) / tot, 100.0 * int(r[iE]) / totE
        if s >= min_pct or e >= min_pct:
            print("%5s  stall %5.1f%%  instr %5.1f%%  %s" % (r[0], s, e, r[1].strip()[:120]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.7)
