"""Summarise ncu CSV output.
  launches: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`
            -> per-kernel launches / total / average / share              (mode: launches)
  full:     `ncu -i rep.ncu-rep --page raw --csv > X.csv` of a --set full capture
            -> one row per launch with the columns the roofline discussion uses (mode: full)
usage: python scripts/ncu_launch_summary.py launches|full in.csv out.csv ["header comment"]
"""
import csv
import re
import sys
from collections import OrderedDict

FULL_COLS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "sm__inst_executed.avg.per_cycle_elapsed", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name) if "<" not in name else name[: name.rfind(">") + 1] if ">(" in name else name


def read_rows(path):
    lines = [l for l in open(path, errors="replace") if not l.startswith("==")]
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"') or l.startswith("ID,"))
    return list(csv.reader(lines[start:]))


def launches(inp, out, note):
    rows = read_rows(inp)
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
        k = short(r[ix["Kernel Name"]])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        if note:
            f.write("# " + note + "\n")
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.1f,%.3f\n' % (k, n, t, t / n, t / tot))


def full(inp, out, note):
    rows = read_rows(inp)
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [c for c in FULL_COLS if c in ix]
    with open(out, "w") as f:
        w = csv.writer(f)
        if note:
            f.write("# " + note + "\n")
        w.writerow(["Kernel Name"] + cols)
        w.writerow([""] + [units[ix[c]] for c in cols])
        for r in rows[2:]:
            if len(r) >= len(hdr):
                w.writerow([r[ix["Kernel Name"]]] + [r[ix[c]] for c in cols])


if __name__ == "__main__":
    mode, inp, out = sys.argv[1:4]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    (launches if mode == "launches" else full)(inp, out, note)
