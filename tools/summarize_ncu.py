#!/usr/bin/env python
"""Turns one GPU visit's ncu output (tools/gpu_round.sh) into the files committed under profiles/:

  python tools/summarize_ncu.py <tag> <round-name>
      gpurun_out/<tag>_launches.csv   (ncu --metrics gpu__time_duration.sum launch list of `bench.py --steps 2`)
      gpurun_out/<tag>_prof_raw.csv   (`--page raw --csv` of one `ncu --set full` capture)
      gpurun_out/<tag>_bench.json     (the bench line of the same visit, NOT run under ncu)
  ->  profiles/<round-name>_launches.csv, _ncu_full_raw.csv, _bench.json, _summary.md, _traffic.json
"""
import csv
import json
import os
import re
import shutil
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = re.compile(r"conv1|dwpw|dw_kernel|pw_gemm|vlad|scores|topk|finalize|merge_lists|split_|dls_|pnp_|icp_|copy_models|f64_to_f32|hamming|gms_")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"<unnamed>::", "", name)
    return re.sub(r"\(.*$", "", name)


def read_launches(path):
    """long-format csv: one row per (launch, metric)."""
    out = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        out.append((int(r["ID"]), short(r["Kernel Name"]), us))
    return out


def last_step(launches):
    """Our kernels of the LAST timed step: the launches between the last two conv1 launches of the timed loop.
    bench.py runs warmup + steps device steps back to back, then the stage timers; a step starts at the PnP setup
    kernel (side stream) or conv1."""
    ours = [(i, n, t) for i, n, t in launches if OURS.search(n)]
    # every step begins with the verifier's first kernel (side stream, issued before conv1); bench.py also solves
    # one PnP batch during set-up, hence the +1 below
    starts = [k for k, (_, n, _) in enumerate(ours) if n.startswith("dls_setup")]
    return ours, starts


def main():
    tag, rnd = sys.argv[1], sys.argv[2]
    src = os.path.join(ROOT, "gpurun_out")
    dst = os.path.join(ROOT, "profiles")
    os.makedirs(dst, exist_ok=True)
    launches = read_launches(os.path.join(src, tag + "_launches.csv"))
    ours, starts = last_step(launches)
    bench = json.loads(open(os.path.join(src, tag + "_bench.json")).read().strip().splitlines()[-1])
    warm, steps = 3, 2
    # step k = kernels from conv1 launch k to conv1 launch k+1, rotated so that the PnP kernels issued just before conv1
    # (side stream) stay with their step: count per-kernel launches over steps `warm .. warm+steps-1` and divide.
    a, b = starts[warm + 1], starts[warm + steps]
    while not ours[b][1].startswith("finalize"):  # the last timed step ends with its finalize launch
        b += 1
    b += 1
    per = OrderedDict()
    for _, n, t in ours[a:b]:
        c = per.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += t
    total = sum(v[1] for v in per.values())
    md = []
    md.append("# %s -- ncu evidence (B200, `bench.py --steps 2 --warmup 3 --no-cpu-baseline`, %d keyframes/step)\n" % (rnd, bench["config"]["keyframes_per_step_per_gpu"]))
    md.append("Produced by `tools/gpu_round.sh %s` on one gpurun box and summarised by `tools/summarize_ncu.py`; numbers printed by a run under ncu are never bench values.\n" % tag)
    md.append("```\nncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n"
              "ncu --set full --clock-control none --import-source on -k regex:<our kernels> -s 40 -c 34 -o prof python bench.py --steps 1 --warmup 3 --no-cpu-baseline\n"
              "ncu -i prof.ncu-rep --page raw --csv > prof_raw.csv      # the .ncu-rep itself exceeds the transfer limit\n```\n")
    md.append("Files: `%s_launches.csv` (every launch, serialised cold-cache times), `%s_ncu_full_raw.csv` (`--set full` raw page), `%s_bench.json` (the bench line of the same visit, not under ncu).\n" % (rnd, rnd, rnd))
    md.append("## Launch list of the two timed steps (per-step averages; shares, not absolutes)\n")
    md.append("| kernel | launches/step | us/step | share |\n|---|---:|---:|---:|")
    stage = {"descriptor": 0.0, "search": 0.0, "pnp": 0.0}
    n_launch = 0
    for n, (c, t) in per.items():
        md.append("| `%s` | %g | %.1f | %.1f %% |" % (n, c / steps, t / steps, 100 * t / total))
        n_launch += c
        key = "pnp" if re.search(r"dls_|pnp_|icp_|copy_models", n) else ("search" if re.search(r"scores|topk|finalize|merge|split_", n) else "descriptor")
        stage[key] += t / steps
    md.append("| total | %g | %.1f | 100 %% |\n" % (n_launch / steps, total / steps))
    st = bench.get("stages_ms", {})
    md.append("Stage shares under ncu: descriptor %.0f %%, search %.0f %%, PnP %.0f %% -- bench.py's CUDA-event stage times in the same visit: %.2f / %.2f / %.2f ms (%.0f / %.0f / %.0f %%); launches per step counted here = %g, bench.py's `gpu_launches` claim = %s.\n" % (
        100 * stage["descriptor"] / (total / steps), 100 * stage["search"] / (total / steps), 100 * stage["pnp"] / (total / steps),
        st.get("descriptor", 0), st.get("search", 0), st.get("pnp", 0),
        100 * st.get("descriptor", 0) / max(sum(st.values()), 1e-9), 100 * st.get("search", 0) / max(sum(st.values()), 1e-9), 100 * st.get("pnp", 0) / max(sum(st.values()), 1e-9),
        n_launch / steps, bench.get("gpu_launches")))

    # ---- full-set raw page (wide format: one row per launch, row 1 = units)
    rows = list(csv.reader(open(os.path.join(src, tag + "_prof_raw.csv"))))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def get(r, name, want_unit=None):
        i = col.get(name)
        if i is None or r[i] == "":
            return float("nan")
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            return float("nan")
        u = units[i]
        if want_unit == "MB":
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}.get(u, 1.0)
        if want_unit == "ms":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
        return v

    md.append("## `--set full`, first captured instance of each kernel\n")
    md.append("| kernel | ms | DRAM read MB | DRAM write MB | DRAM % | L2 (lts) % | SM % | issue % | warps active % | regs | fp64 pipe % | tensor pipe % |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    seen = OrderedDict()
    for r in data:
        n = short(r[col["Kernel Name"]])
        if n in seen:
            continue
        seen[n] = r
        md.append("| `%s` | %.3f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.1f | %.1f |" % (
            n, get(r, "gpu__time_duration.sum", "ms"), get(r, "dram__bytes_read.sum", "MB"), get(r, "dram__bytes_write.sum", "MB"),
            get(r, "dram__bytes_read.sum.pct_of_peak_sustained_elapsed") + get(r, "dram__bytes_write.sum.pct_of_peak_sustained_elapsed"), get(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            get(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            get(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"), get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            int(get(r, "launch__registers_per_thread")), get(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")))
    md.append("")
    traffic = {}
    for n, r in seen.items():
        if n.startswith("scores"):
            v = (get(r, "dram__bytes_read.sum", "MB") + get(r, "dram__bytes_write.sum", "MB")) * 1e6
            traffic[n + "_dram_bytes_per_launch"] = v
            traffic.setdefault(re.sub(r"<.*$", "", n) + "_dram_bytes_per_launch", v)  # template arguments stripped
    json.dump(traffic, open(os.path.join(dst, rnd + "_traffic.json"), "w"), indent=1)
    shutil.copy(os.path.join(src, tag + "_launches.csv"), os.path.join(dst, rnd + "_launches.csv"))
    shutil.copy(os.path.join(src, tag + "_prof_raw.csv"), os.path.join(dst, rnd + "_ncu_full_raw.csv"))
    json.dump(bench, open(os.path.join(dst, rnd + "_bench.json"), "w"), indent=1)
    notes = os.path.join(dst, rnd + "_reading.md")
    if os.path.exists(notes):
        md.append(open(notes).read())
    open(os.path.join(dst, rnd + "_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
