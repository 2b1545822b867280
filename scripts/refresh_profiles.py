#!/usr/bin/env python3
"""Turn the outputs of scripts/gpu_profile.sh (gpurun_out/*<tag>*) into the tracked evidence under profiles/.
usage: scripts/refresh_profiles.py <tag>      (run in the container, after the gpurun call has merged its files)"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

shutil.copy(os.path.join(G, "bench_%s_k1.json" % tag), os.path.join(P, "r1_bench_k1.json"))
shutil.copy(os.path.join(G, "bench_%s_ref.json" % tag), os.path.join(P, "r1_bench_reference_arm.json"))
shutil.copy(os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "r1_launches_bench_k1.csv"))

# ---- launch shares
rows = list(csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[start + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        launches.append((r[ki].split("(")[0].replace("void ", ""), v / 1000.0 if r[ui] == "ns" else v))
steady = launches[len(launches) // 3:]                       # skip allocation / warm-up launches
agg = collections.OrderedDict()
for k, us in steady:
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
step_kernels = ("k_fast", "k_octree", "k_resize", "k_orient", "k_blur", "k_repack", "k_pyramid")
is_step = lambda k: any(t in k for t in step_kernels)
tot = sum(a[1] for k, a in agg.items() if is_step(k))
with open(os.path.join(P, "r1_launch_shares.txt"), "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline` (round 1, current kernels)\n"
            "# gpu__time_duration.sum per launch, --clock-control none; cold-cache and serialised: compare SHARES, not absolutes\n"
            "# last two thirds of the %d captured launches; source: r1_launches_bench_k1.csv\n\n" % len(launches))
    f.write("# kernels of the extract+describe step (shares of the step)\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if is_step(k):
            f.write("%-32s launches %4d  total %9.1f us  avg %8.1f us  share %5.1f%%\n" % (k, a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
    f.write("\n# other launches of the same run (the 5000x5000 matcher timing loop of bench.py, torch fills)\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if not is_step(k):
            f.write("%-32s launches %4d  total %9.1f us  avg %8.1f us\n" % (k[:32], a[0], a[1], a[1] / a[0]))

# ---- full captures
traffic = {}
for rep, out in (("prof_%s_main.ncu-rep" % tag, "r1_ncu_full_main_kernels.txt"), ("prof_%s_resize.ncu-rep" % tag, "r1_ncu_full_resize_levels.txt")):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), path], capture_output=True, text=True).stdout
    hot = ""
    for kre in ("k_fast_fused", "k_blur", "k_orient_desc", "k_octree", "k_resize"):
        h = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_source_hot.py"), path, kre, "14"], capture_output=True, text=True)
        if h.returncode == 0 and h.stdout.strip():
            hot += "\n-- hottest source lines, %s\n%s" % (kre, h.stdout)
    open(os.path.join(P, out), "w").write("# ncu --set full --clock-control none --import-source on, `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n" + txt + hot)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h = rr[0]
    kk, a, b = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    units = rr[1]
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rr[2:]:
        name = r[kk].split("(")[0].replace("void ", "").split("<")[0]
        val = float(r[a]) * mult[units[a]] + float(r[b]) * mult[units[b]]
        traffic.setdefault(name, []).append(val)
# ---- one table: per kernel duration, DRAM traffic and throughput, pipe utilisation (north_star: HBM GB/s + integer pipe)
want = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", "regs")]
lines = ["# per-kernel summary of the `ncu --set full` captures (one mid capture per kernel; batch of 32 frames, 1242x375)",
         "# DRAM GB/s = (dram read + write) / duration; peak copy bandwidth on this pod 6545 GB/s (MEASURED_PEAKS.json); alu% = integer ALU pipe",
         "%-28s %9s %9s %9s %7s %6s %6s %7s %6s %10s %5s" % ("kernel", "us", "dram MB", "DRAM GB/s", "dram%", "alu%", "fma%", "issue%", "occ%", "warp-inst", "regs")]
for rep in ("prof_%s_main.ncu-rep" % tag, "prof_%s_resize.ncu-rep" % tag):
    path = os.path.join(G, rep)
    if not os.path.exists(path):
        continue
    rr = list(csv.reader(subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, units = rr[0], rr[1]
    kk = h.index("Kernel Name")
    groups = collections.OrderedDict()
    for r in rr[2:]:
        groups.setdefault(r[kk].split("(")[0].replace("void ", ""), []).append(r)
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    for name, rs in groups.items():
        picks = rs if name.startswith("k_resize") else [rs[len(rs) // 2]]
        for r in picks:
            v = {}
            for m, short in want:
                if m in h:
                    i = h.index(m)
                    v[short] = float(r[i].replace(",", "")) * mult.get(units[i], 1.0)
            mb = (v.get("rd", 0) + v.get("wr", 0)) / 1e6
            lines.append("%-28s %9.1f %9.2f %9.0f %7.1f %6.1f %6.1f %7.1f %6.1f %10.0f %5.0f" % (
                name[:28], v.get("us", 0), mb, mb * 1e6 / (v.get("us", 1) * 1e-6) / 1e9, v.get("dram%", 0), v.get("alu%", 0), v.get("fma%", 0),
                v.get("issue%", 0), v.get("occ%", 0), v.get("inst", 0), v.get("regs", 0)))
open(os.path.join(P, "r1_kernel_table.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))

tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), mean over the captured launches, from `ncu --set full "
                  "--clock-control none` of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (batch of 32 frames, 1242x375); "
                  "k_resize_tma is the SUM over the 7 level launches; sources: profiles/r1_ncu_full_*.txt",
      "k1": {}}
for k, v in traffic.items():
    tj["k1"][k] = int(sum(v) / len(v) * (7 if k.startswith("k_resize") else 1))
json.dump(tj, open(os.path.join(P, "r1_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, "r1_launch_shares.txt")).read())
print(json.dumps(tj["k1"]))
