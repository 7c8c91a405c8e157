"""Reads `ncu --page source --csv` output of one kernel and prints the instructions with the most stall samples."""
import csv
import sys


def main(path, ntop=45):
    rows = list(csv.reader(open(path)))
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # n-th kernel section of the file
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    rows = rows[starts[which] : (starts[which + 1] if which + 1 < len(starts) else len(rows))]
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    tot = sum(int(r[iS]) for r in data)
    print(rows[0][1][:100])
    print("total samples", tot, "instructions", len(data))
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    agg = {}
    for r in data:
        for i in stall_cols:
            agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
    print("stalls:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    top = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:ntop]
    for k in sorted(top):
        r = data[k]
        st = sorted([(int(r[i]), hdr[i][6:]) for i in stall_cols], reverse=True)[:2]
        print("%5d %6d %5.1f%%  ex=%8s  %-64s %s" % (k, int(r[iS]), 100 * int(r[iS]) / tot, r[iEx], r[iSrc].strip()[:64], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
