"""Turns ncu CSV output into the short markdown summaries kept under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_<tag>.csv  > profiles/<tag>_launches.md
    ncu -i gpurun_out/prof_<tag>.ncu-rep --page raw --csv > gpurun_out/prof_<tag>_raw.csv
    python scripts/summarize_ncu.py full gpurun_out/prof_<tag>_raw.csv      > profiles/<tag>_full.md

`launches`: one row per kernel name with launch count, total and mean gpu__time_duration and the share of the
summed device time (ncu serialises launches and runs them cold-cache: compare shares, not absolutes).
`full`: the roofline-relevant metrics of every profiled launch of an `ncu --set full` capture.
"""
import csv
import re
import sys
from collections import OrderedDict

KEY_METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
]


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name.strip())
    if name.endswith(")"):                     # drop the argument list (balanced from the end)
        depth = 0
        for i in range(len(name) - 1, -1, -1):
            depth += name[i] == ")"
            depth -= name[i] == "("
            if depth == 0:
                name = name[:i]
                break
    name = re.sub(r"^(\w+::)*(<unnamed>|\(anonymous namespace\))::", "", name)
    return name[:90]


def launches(path):
    rows = OrderedDict()
    total = 0.0
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        k = short(r["Kernel Name"])
        n, t = rows.get(k, (0, 0.0))
        rows[k] = (n + 1, t + ns)
        total += ns
    print(f"| kernel | launches | total us | mean us | share of device time |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {100 * t / total:.1f} % |")
    print(f"| **all** | {sum(n for n, _ in rows.values())} | {total / 1e3:.1f} | | 100 % |")


def full(path):
    with open(path, newline="") as fh:
        lines = [l for l in fh if not l.startswith("==")]
    rd = csv.reader(lines)
    header = next(rd)
    units = next(rd)
    idx = {h: i for i, h in enumerate(header)}
    for row in rd:
        if len(row) < len(header):
            continue
        print(f"### `{short(row[idx['Kernel Name']])}`  grid {row[idx['Grid Size']]} block {row[idx['Block Size']]}\n")
        print("| metric | value | unit |")
        print("|---|---:|---|")
        for m in KEY_METRICS:
            if m in idx:
                print(f"| {m} | {row[idx[m]]} | {units[idx[m]]} |")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
