#!/usr/bin/env python
"""Turns the ncu artefacts in gpurun_out/ into the text summaries committed under profiles/.
usage: python profiles/summarize.py <tag>   (reads gpurun_out/launches_<tag>.csv, prof_*_<tag>.ncu-rep, prof_tapes.log)"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1g"
out = io.StringIO()

# ---- launch list: every kernel with its device time and the shape of its tape ----
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 5]
hdr = rows[0]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
tapes = [l for l in open(os.path.join(G, "prof_tapes.log")) if l.startswith("[fmc tape]")]
ti = 0
out.write(f"# launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none, benchmarks/profile_cases.py\n")
out.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
out.write(f"{'#':>4} {'kernel':28s} {'grid':>6} {'us':>9}  tape: n instr ptrs leaves stores ring regs ctas/sm | GB/s algorithmic, GB/s touched\n")
tot = {}
for k, r in enumerate(rows[1:]):
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("fmc::", "").replace("<unnamed>::", "")[:28]
    us = float(r[vi].replace(",", "")) / 1000.0
    grid = r[gi].strip("()").split(",")[0]
    extra = ""
    if "tape_kernel" in name and ti < len(tapes):
        m = re.search(r"n=(\d+) instr=(\d+).*ptrs=(\d+) leaves=(\d+) stores=(\d+) ring=(\d+) regs=(\d+).*ctas/sm=(\d+)", tapes[ti]); ti += 1
        n, instr, ptrs, leaves, stores, ring, regs, cps = map(int, m.groups())
        extra = f"{n} {instr} {ptrs} {leaves} {stores} {ring} {regs} {cps} | {4 * n * (leaves + stores) / us / 1e3:7.0f} {4 * n * ptrs / us / 1e3:7.0f}"
    tot[name] = tot.get(name, 0.0) + us
    out.write(f"{k:4d} {name:28s} {grid:>6} {us:9.1f}  {extra}\n")
out.write("\n# time per kernel name (us, share)\n")
s = sum(tot.values())
for name, us in sorted(tot.items(), key=lambda kv: -kv[1]):
    out.write(f"{name:28s} {us:10.1f} {100 * us / s:5.1f}%\n")
open(os.path.join(ROOT, "profiles", f"launches_{tag}.txt"), "w").write(out.getvalue())

# ---- full captures: the metrics the roofline numbers come from ----
WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]
for rep in sorted(f for f in os.listdir(G) if f.endswith(f"_{tag}.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    if len(rr) < 3:
        continue
    h, units = rr[0], rr[1]
    o = io.StringIO()
    o.write(f"# {rep}: ncu --set full --clock-control none --import-source on (selected metrics; units from ncu)\n")
    stall = [i for i, x in enumerate(h) if "issue_stalled" in x and x.endswith("per_issue_active.ratio")]
    for r in rr[2:]:
        o.write("\n")
        for w in WANT:
            if w in h:
                i = h.index(w)
                o.write(f"{w:90s} {r[i]:>18s} {units[i]}\n")
        top = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, h[i]) for i in stall), reverse=True)[:6]
        o.write("top stalls (warps per issue-active cycle): " + ", ".join(f"{n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, n in top) + "\n")
    open(os.path.join(ROOT, "profiles", rep.replace(".ncu-rep", ".txt")), "w").write(o.getvalue())
print("written:", sorted(os.listdir(os.path.join(ROOT, "profiles"))))
