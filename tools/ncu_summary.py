#!/usr/bin/env python
"""Per-kernel table from an .ncu-rep (ncu -i REP --page raw --csv): time, DRAM bytes, pipe / memory utilisation.
Kernels launched several times are averaged; prints GitHub markdown.

    python tools/ncu_summary.py gpurun_out/prof_fr.ncu-rep [more.ncu-rep ...]
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

COLS = [("gpu__time_duration.sum", "time [us]", 1e-3),
        ("dram__bytes_read.sum", "dram rd [MB]", 1e-6),
        ("dram__bytes_write.sum", "dram wr [MB]", 1e-6),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram %", 1),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
        ("launch__registers_per_thread", "regs", 1),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex %", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts %", 1),
        ("sm__inst_executed.avg.pct_of_peak_sustained_active", "issue %", 1)]


def table(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    kn = idx["Kernel Name"]
    acc = OrderedDict()
    for r in rows[2:]:
        name = r[kn].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = acc.setdefault(name, {"n": 0, "v": [0.0] * len(COLS)})
        a["n"] += 1
        for j, (m, _, sc) in enumerate(COLS):
            if m in idx and r[idx[m]] not in ("", "n/a"):
                v = float(r[idx[m]].replace(",", ""))
                u = units[idx[m]]
                if m.startswith("gpu__time") and u in ("us", "usecond"):
                    v *= 1e3
                if m.startswith("gpu__time") and u in ("ms", "msecond"):
                    v *= 1e6
                if m.startswith("dram__bytes"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                a["v"][j] += v * sc
    print(f"### {rep}\n")
    print("| kernel | launches | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    for name, a in acc.items():
        print(f"| {name} | {a['n']} | " + " | ".join(f"{v / a['n']:.1f}" for v in a["v"]) + " |")
    print()


if __name__ == "__main__":
    for rep in sys.argv[1:]:
        table(rep)
