#!/usr/bin/env python3
"""Turn the outputs of scripts/gpu_profile.sh (gpurun_out/*<tag>*) into the tracked evidence under profiles/<tag>_*.
usage: scripts/refresh_profiles.py <tag>      (run in the container, after the gpurun call has merged its files)"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

for src, dst in (("bench_%s_k1.json", "%s_bench_k1.json"), ("bench_%s_k1_20steps.json", "%s_bench_k1_20steps.json"), ("bench_%s_ref.json", "%s_bench_reference_arm.json"),
                 ("bench_%s_k2.json", "%s_bench_k2.json"), ("bench_%s_k4.json", "%s_bench_k4.json"), ("launches_%s.csv", "%s_launches_bench_k1.csv")):
    if os.path.exists(os.path.join(G, src % tag)):
        shutil.copy(os.path.join(G, src % tag), os.path.join(P, dst % tag))

# ---- launch shares of the default bench command (scripts/gpu_final.sh runs, e.g. tag r2b, have no launch list: skipped)
have_launches = os.path.exists(os.path.join(G, "launches_%s.csv" % tag))
rows = list(csv.reader(open(os.path.join(G, "launches_%s.csv" % tag)))) if have_launches else [["Kernel Name", "Metric Value", "Metric Unit"]]
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[start + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        launches.append((r[ki].split("(")[0].replace("void ", ""), v / 1000.0 if r[ui] == "ns" else v))
steady = launches[len(launches) // 3:]                       # skip allocation / priming launches
agg = collections.OrderedDict()
for k, us in steady:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
step_kernels = ("k_fast", "k_octree", "k_resize", "k_orient", "k_blur", "k_repack", "k_pyramid")
is_step = lambda k: any(t in k for t in step_kernels)
tot = sum(a[1] for k, a in agg.items() if is_step(k)) or 1.0
with open(os.path.join(P, "%s_launch_shares.txt" % tag) if have_launches else os.devnull, "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras` (round 2 kernels)\n"
            "# gpu__time_duration.sum per launch, --clock-control none; cold-cache and serialised: compare SHARES, not absolutes\n"
            "# last two thirds of the %d captured launches; source: %s_launches_bench_k1.csv\n\n" % (len(launches), tag))
    f.write("# kernels of the extract+describe step (shares of the step)\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if is_step(k):
            f.write("%-32s launches %4d  total %9.1f us  avg %8.1f us  share %5.1f%%\n" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    f.write("\n# other launches of the same run\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if not is_step(k):
            f.write("%-32s launches %4d  total %9.1f us  avg %8.1f us\n" % (k[:32], a[0], a[1], a[1] / a[0]))

# ---- one table per workload from the raw metric pages of the `ncu --set full` captures
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("lts__t_sector_hit_rate.pct", "l2hit%")]
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
frames = {"k1": "batch of 32 frames, 1242x375, 2000 features", "k2": "batch of 64 frames, 1920x1080, 5000 features",
          "k4": "batch of 8 frames, 3840x2160, 12 levels, 10000 features", "matcher": "5000 x 5000 descriptors",
          "next": "the SURVEY 8f kernels on K1-sized inputs (scripts/run_next_rows.py; first capture of every kernel)"}
traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes) from `ncu --set full --clock-control none` captures of "
                       "scripts/run_workload.py <workload> (one batch; k_resize_sep is the SUM over the level launches); table: profiles/%s_kernel_table.txt" % tag}
lines = []
for wl in ("k1", "k2", "k4", "matcher", "next"):
    path = os.path.join(G, "ncu_raw_%s_%s.csv" % (tag, wl))
    if not os.path.exists(path):
        continue
    rr = list(csv.reader(open(path)))
    h, units = rr[0], rr[1]
    kk = h.index("Kernel Name")
    lines += ["", "# %s: %s -- one capture per launch of one batch (`ncu --set full --clock-control none`), DRAM GB/s = (read + write) / duration" % (wl, frames[wl]),
              "%-30s %8s %9s %9s %6s %6s %6s %6s %6s %7s %6s %10s %5s %7s" % ("kernel", "us", "dram MB", "DRAM GB/s", "dram%", "l2hit%", "alu%", "fma%", "xu%", "issue%", "occ%", "warp-inst", "regs", "grid")]
    tw = {}
    total_us = 0.0
    seen_next = set()
    for r in rr[2:]:
        name = r[kk].split("(")[0].replace("void ", "")
        if wl == "next":
            if name in seen_next:
                continue
            seen_next.add(name)
        v = {}
        for m, short in want:
            if m in h:
                i = h.index(m)
                try:
                    v[short] = float(r[i].replace(",", "")) * mult.get(units[i], 1.0)
                except ValueError:
                    v[short] = 0.0
        mb = (v.get("rd", 0) + v.get("wr", 0)) / 1e6
        total_us += v.get("us", 0)
        lines.append("%-30s %8.1f %9.2f %9.0f %6.1f %6.1f %6.1f %6.1f %6.1f %7.1f %6.1f %10.0f %5.0f %7.0f" % (
            name[:30], v.get("us", 0), mb, mb * 1e6 / max(v.get("us", 1) * 1e-6, 1e-12) / 1e9, v.get("dram%", 0), v.get("l2hit%", 0), v.get("alu%", 0), v.get("fma%", 0), v.get("xu%", 0),
            v.get("issue%", 0), v.get("occ%", 0), v.get("inst", 0), v.get("regs", 0), v.get("grid", 0)))
        base = name.split("<")[0]
        tw[base] = tw.get(base, 0) + int(v.get("rd", 0) + v.get("wr", 0))
    lines.append("# sum of the captured launches: %.1f us (serialised, cold caches)" % total_us)
    traffic[wl] = tw
    hot = os.path.join(G, "ncu_hot_%s_%s.txt" % (tag, wl))
    if os.path.exists(hot):
        shutil.copy(hot, os.path.join(P, "%s_ncu_hot_lines_%s.txt" % (tag, wl)))
open(os.path.join(P, "%s_kernel_table.txt" % tag), "w").write(
    "# per-kernel summary of the round-2 `ncu --set full` captures (scripts/gpu_profile.sh); peak copy bandwidth on this pod 6545 GB/s (MEASURED_PEAKS.json);\n"
    "# l2hit% = lts__t_sector_hit_rate; alu% = integer ALU pipe, fma% = FMA pipe (IMAD, IDP), xu% = XU pipe (POPC), all as % of peak while the SM was active" + "\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(P, "%s_traffic.json" % tag), "w"), indent=1)
print(open(os.path.join(P, "%s_kernel_table.txt" % tag)).read())
if have_launches:
    print(open(os.path.join(P, "%s_launch_shares.txt" % tag)).read())
