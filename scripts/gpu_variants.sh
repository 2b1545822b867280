#!/bin/bash
# K1 device-resident loop on several builds of the library, interleaved twice.  usage: scripts/gpu_variants.sh lib1.so lib2.so ...
set -u
mkdir -p gpurun_out
S=gpurun_out/variants_summary.txt; : > $S
for round in 1 2; do
  for lib in "$@"; do
    tag=$(basename $lib .so)
    ORBX_LIBRARY=$lib timeout 150 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-extras > gpurun_out/var_${tag}_$round.json 2> gpurun_out/var_${tag}_$round.err
    python - <<P >> $S
import json
try:
    d = json.loads(open("gpurun_out/var_${tag}_$round.json").read().strip().splitlines()[-1])
    print("$tag $round value %.0f ms %.4f fast %.4f parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["stage_ms"]["fast"], d.get("parity_checked")))
except Exception as e:
    print("$tag $round failed", e)
P
  done
done
cat $S
