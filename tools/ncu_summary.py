#!/usr/bin/env python
"""Turns the ncu outputs of one profiling call into the tracked summaries under profiles/.

    python tools/ncu_summary.py <launches.csv> <report.ncu-rep> <round tag> <steps in the launch list>

launches.csv  : ncu --metrics gpu__time_duration.sum --csv --log-file ...   (every launch of a bench run)
report.ncu-rep: ncu --set full --import-source on -k regex:...               (one launch per hot kernel)
Writes profiles/<tag>_launches_ncu.csv (copy), profiles/<tag>_kernels.json, profiles/<tag>_screen_traffic.json
and prints the markdown tables used in profiles/<tag>_summary.md."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launches, report, tag, nsteps = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
prof = os.path.join(ROOT, "profiles")
shutil.copy(launches, os.path.join(prof, f"{tag}_launches_ncu.csv"))

rows = list(csv.DictReader(l for l in open(launches) if not l.startswith("==")))
per = len(rows) // nsteps
agg, tot = collections.OrderedDict(), 0.0
for r in rows[(nsteps - 1) * per:]:
    name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r["Kernel Name"])).replace("<unnamed>::", "")
    v, u = float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
    ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ms; tot += ms
print("| kernel | launches / step | ms / step (ncu: serialised, cold cache) | share |\n|---|---|---|---|")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"| `{k}` | {c} | {ms:.3f} | {100 * ms / tot:.1f}% |")
print(f"| **total** | {sum(c for c, _ in agg.values())} | {tot:.3f} | 100% |\n")

raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
out = {}
for r in rr[2:]:
    m = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", m["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
    out[name] = {k: (m[k], units[hdr.index(k)]) for k in want if k in m}
json.dump(out, open(os.path.join(prof, f"{tag}_kernels.json"), "w"), indent=1)
for name, m in out.items():
    print(f"### `{name}`\n\n| metric | value |\n|---|---|")
    for k, (v, u) in m.items():
        print(f"| `{k}` | {v} {u} |")
    print()
    if "knn_screen_pair" in name:
        gb = lambda k: float(m[k][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12, "Kbyte": 1e3, "byte": 1.0}[m[k][1]]
        json.dump({"kernel": name, "workload": "C2 1M x 384 cosine k=16, 1 GPU, one launch", "dram_bytes_read": gb("dram__bytes_read.sum"),
                   "dram_bytes_write": gb("dram__bytes_write.sum"), "traffic_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                   "source": f"ncu --set full --clock-control none, profiles/{tag}_kernels.json"}, open(os.path.join(prof, f"{tag}_screen_traffic.json"), "w"), indent=1)
