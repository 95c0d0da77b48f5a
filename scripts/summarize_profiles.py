#!/usr/bin/env python
"""Turns the raw ncu artefacts in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r01          # round tag
Reads gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum launch list) and gpurun_out/prof_*.ncu-rep
(ncu --set full captures) and writes profiles/<tag>_launches.md and profiles/<tag>_<name>_full.md.
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "profiles"
G = ROOT / "gpurun_out"

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def short(name: str) -> str:
    if "vita::" in name:
        return name.split("vita::", 1)[1].split("(")[0]
    return "torch:" + name.replace("void ", "")[:60]


def launches(tag):
    src = G / "launches.csv"
    if not src.exists():
        return
    lines = src.read_text().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(lines[start:]))
    agg = collections.OrderedDict()
    for r in rows:
        d = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)",
           "", f"source: gpurun_out/launches.csv, {len(rows)} launches, total {tot / 1e6:.2f} ms", "",
           "| kernel | launches | total us | us/launch | share |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / v[0] / 1e3:.2f} | {100 * v[1] / tot:.1f}% |")
    (OUT / f"{tag}_launches.md").write_text("\n".join(out) + "\n")
    print("wrote", OUT / f"{tag}_launches.md")


def full(tag):
    for rep in sorted(G.glob("prof_*.ncu-rep")):
        res = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(res.stdout.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        out = [f"# {tag}: ncu --set full --clock-control none, {rep.name}", ""]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            out += [f"## `{short(name)}`", "", "| metric | value | unit |", "|---|---:|---|"]
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    out.append(f"| {m} | {r[i]} | {units[i]} |")
            out.append("")
        (OUT / f"{tag}_{rep.stem}_full.md").write_text("\n".join(out) + "\n")
        print("wrote", OUT / f"{tag}_{rep.stem}_full.md")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    OUT.mkdir(exist_ok=True)
    launches(tag)
    full(tag)
