"""`ncu -i rep --page raw --csv` -> a markdown table, one row per kernel name (first instance + launch count + total time).
  python tools/ncu_table.py raw.csv [title] > table.md"""
import csv
import re
import sys
from collections import OrderedDict


def main(path, title="ncu --set full"):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, want=None):
        i = col.get(name)
        if i is None or r[i] == "":
            return float("nan")
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return float("nan")
        u = units[i]
        if want == "MB":
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
        if want == "us":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        return v

    seen = OrderedDict()
    for r in data:
        n = re.sub(r"\(.*$", "", r[col["Kernel Name"]].replace("void ", "").replace("<unnamed>::", ""))
        e = seen.setdefault(n, [r, 0, 0.0])
        e[1] += 1
        e[2] += get(r, "gpu__time_duration.sum", "us")
    print("## %s\n" % title)
    print("| kernel | launches | us (first) | us (all) | DRAM r MB | DRAM w MB | DRAM % | L2 % | SM % | issue % | warps active % | regs |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for n, (r, c, t) in seen.items():
        print("| `%s` | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d |" % (
            n, c, get(r, "gpu__time_duration.sum", "us"), t, get(r, "dram__bytes_read.sum", "MB"), get(r, "dram__bytes_write.sum", "MB"),
            get(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + get(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"),
            get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            get(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            int(get(r, "launch__registers_per_thread"))))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "ncu --set full")
