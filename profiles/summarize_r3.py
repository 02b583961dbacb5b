#!/usr/bin/env python
"""Round-2 (final state: windows of time steps) summaries of the ncu artefacts in gpurun_out/ (scratch) -> profiles/ (tracked).
usage: python profiles/summarize_r3.py <tag> [report.ncu-rep ...]
  gpurun_out/launches_<tag>.csv : `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv` of
                                  `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-calibration --no-extras`
  reports                        : `ncu --set full --clock-control none --import-source on` captures (scripts/gpu_evidence_r3.sh)"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1]
reports = sys.argv[2:]

path = os.path.join(G, f"launches_{tag}.csv")
if os.path.exists(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    out = io.StringIO()
    out.write(f"# launch list ({tag}): ncu --metrics gpu__time_duration.sum --clock-control none -c 700, python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-calibration --no-extras\n")
    out.write("# per-launch times are cold-cache and serialised: compare SHARES with the bench line, not absolutes.\n")
    out.write("# One LMM step = the store-only launches of the simulation (tape_kernel<0>: windows of 3 Euler time steps) + 144 chain->reduce launches (tape_kernel<1>: one per swaption).\n")
    names = []
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("fmc::", "").replace("<unnamed>::", "")
        names.append((name, float(r[vi].replace(",", "")) / 1000.0, r[gi].strip("()").split(",")[0]))
    tape = [i for i, (n, _, _) in enumerate(names) if "tape_kernel" in n]
    # a step = a run of store-only launches (the windows of the simulation) followed by the 144 chain->reduce launches of the swaptions
    runs = []
    for i in tape:
        kind = 0 if "<0" in names[i][0] else 1
        if runs and runs[-1][0] == kind: runs[-1][1].append(i)
        else: runs.append([kind, [i]])
    steps = [runs[k][1] + runs[k + 1][1] for k in range(len(runs) - 1) if runs[k][0] == 0 and runs[k + 1][0] == 1 and len(runs[k + 1][1]) == 144]
    step = steps[1] if len(steps) > 1 else steps[0]          # the first timed step (the warm-up step comes before it)
    out.write(f"\n# second LMM step of the run (launches {step[0]}..{step[-1]} of the list): {sum(1 for i in step if '<0' in names[i][0])} window launches + 144 swaption launches\n")
    cls = {}
    for i in step:
        n, us, g = names[i]
        key = "simulation windows tape_kernel<0>" if "<0" in n else "swaption kernels tape_kernel<1>"
        c = cls.setdefault(key, [0, 0.0]); c[0] += 1; c[1] += us
    tot = sum(c[1] for c in cls.values())
    for k, (cnt, us) in cls.items():
        out.write(f"{k:40s} {cnt:4d} launches {us / 1000:8.3f} ms {100 * us / tot:5.1f}% of the step's kernel time\n")
    out.write(f"{'all interpreter launches of the step':40s} {len(step):4d} launches {tot / 1000:8.3f} ms\n")
    out.write("\n# time per kernel name over the whole list (us, share)\n")
    agg = {}
    for n, us, _ in names:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us
    s = sum(a[1] for a in agg.values())
    for n, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write(f"{n:50s} {cnt:5d} {us:12.1f} {100 * us / s:5.1f}%\n")
    out.write(f"\n{'#':>4} {'kernel':46s} {'grid':>6} {'us':>9}\n")
    for i, (n, us, g) in enumerate(names):
        out.write(f"{i:4d} {n:46s} {g:>6} {us:9.1f}\n")
    open(os.path.join(ROOT, "profiles", f"launches_{tag}.txt"), "w").write(out.getvalue())

WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg"]
for rep in reports:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    if len(rr) < 3:
        continue
    h, units = rr[0], rr[1]
    o = io.StringIO()
    o.write(f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on (selected metrics; units from ncu)\n")
    stall = [i for i, x in enumerate(h) if "issue_stalled" in x and x.endswith("per_issue_active.ratio")]
    for r in rr[2:]:
        o.write("\n")
        for w in WANT:
            if w in h:
                i = h.index(w)
                o.write(f"{w:90s} {r[i]:>18s} {units[i]}\n")
        try:
            rd, wr = float(r[h.index("dram__bytes_read.sum")].replace(",", "")), float(r[h.index("dram__bytes_write.sum")].replace(",", ""))
            ur, uw = units[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_write.sum")]
            us = float(r[h.index("gpu__time_duration.sum")].replace(",", ""))
            if ur == uw == "Mbyte" and units[h.index("gpu__time_duration.sum")] == "us":
                o.write(f"{'dram traffic (read + write) / duration':90s} {(rd + wr) / us:18.2f} TB/s\n")
        except (ValueError, KeyError):
            pass
        top = sorted(((float(r[i].replace(",", "")) if r[i] else 0.0, h[i]) for i in stall), reverse=True)[:6]
        o.write("top stalls (warps per issue-active cycle): " + ", ".join(f"{n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, n in top) + "\n")
    open(os.path.join(ROOT, "profiles", os.path.basename(rep).replace(".ncu-rep", ".txt")), "w").write(o.getvalue())
print("written:", sorted(os.listdir(os.path.join(ROOT, "profiles"))))
