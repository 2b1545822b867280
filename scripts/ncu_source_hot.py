#!/usr/bin/env python3
"""Per-source-line instruction counts / stall samples from an .ncu-rep source page (needs -lineinfo).
usage: scripts/ncu_source_hot.py report.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre, "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the CSV holds one block per kernel result: [File Name..] [header] rows...
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": []}; blocks.append(cur)
    elif cur is not None and r and r[0].isdigit():
        cur["rows"].append(r)
b = blocks[0]
h = b["hdr"]
iline, isrc, iinst, isamp, ithr = 0, 1, h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
agg = collections.OrderedDict()
for r in b["rows"]:
    key = (int(r[iline]), r[isrc].strip())
    a = agg.setdefault(key, [0, 0, 0, 0])
    num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
    a[0] += num(r[iinst]); a[1] += num(r[isamp]); a[2] += num(r[ithr]); a[3] += 1
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("kernel block 0: total warp instructions %d, samples %d, sass lines %d" % (tot_i, tot_s, len(b["rows"])))
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5d  inst %5.1f%%  samp %5.1f%%  sass %3d  thr/inst %4.1f | %s" % (ln, 100 * a[0] / tot_i, 100 * a[1] / max(1, tot_s), a[3], a[2] / max(1, a[0]), src[:110]))
