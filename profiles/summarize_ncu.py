"""Turns ncu captures (gpurun_out/*.ncu-rep, launch-list csv) into the small text summaries committed under profiles/.
usage: python profiles/summarize_ncu.py <tag>      (reads gpurun_out/, writes profiles/<tag>_*.txt/json)"""
import collections
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles"
SRC = ROOT / "gpurun_out"
KEEP = [r"^Kernel Name$", r"^launch__grid_size$", r"^launch__registers_per_thread$", r"^gpu__time_duration\.sum$",
        r"^dram__bytes_read\.sum$", r"^dram__bytes_write\.sum$", r"^dram__bytes_read\.sum\.per_second$",
        r"^dram__bytes_read\.sum\.pct_of_peak_sustained_elapsed$", r"^lts__t_sector_hit_rate\.pct$",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$", r"^sm__inst_executed_pipe_xu\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|barrier|wait|branch_resolving|membar)_per_issue_active\.ratio$",
        r"^smsp__inst_executed\.sum$", r"^sm__cycles_elapsed\.max$", r"^smsp__sass_inst_executed_op_tmem_ldt\.sum$"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summarize_rep(rep, tag):
    rows = raw(rep)
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [i for i, h in enumerate(hdr) if any(re.search(p, h) for p in KEEP)]
    lines = [f"# {rep.name}: ncu --set full --clock-control none (per launch; cold-cache, serialised)"]
    for i in idx:
        lines.append(f"{hdr[i]:85s} [{units[i]}] " + " | ".join(r[i][:48] for r in data))
    (OUT / f"{tag}_{rep.stem}.txt").write_text("\n".join(lines) + "\n")
    return {hdr[i]: [r[i] for r in data] for i in idx}, dict(zip(hdr, units))


def summarize_launches(path, tag):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        name = re.sub(r"\(.*", "", r[ki])[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = [f"# {Path(path).name}: ncu --metrics gpu__time_duration.sum --clock-control none, one profiled step "
             f"(mel+encoder+projector+prefill+2 eager decode steps); {len(data)} launches, {tot/1e3:.1f} ms total",
             "# cold-cache, serialised launches: compare SHARES with the stage split of bench.py, not absolutes",
             f"{'ms':>10s} {'share':>7s} {'launches':>8s}  kernel"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{v[1]/1e3:10.3f} {100*v[1]/tot:6.1f}% {v[0]:8d}  {k}")
    (OUT / f"{tag}_launch_list.txt").write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    for rep in sorted(SRC.glob("*.ncu-rep")):
        summarize_rep(rep, tag)
    for p in SRC.glob("launches_*.csv"):
        summarize_launches(p, tag)
    print("wrote", sorted(x.name for x in OUT.glob(f"{tag}_*")))
