#!/bin/bash
# A/B of two builds of the library on one box: GPU tests on the new one, then the device-resident loops on both.
# usage: scripts/gpu_ab.sh [previous liborbx.so]      (variants: new, prev, new4 = debug build with the FAST CTAs padded to 4 per SM)
set -u
prev=${1:-multimot_track_b200/liborbx_prev.so}
mkdir -p gpurun_out
S=gpurun_out/ab_summary.txt
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?" > $S
tail -3 gpurun_out/ab_tests.log >> $S
one() {  # workload variant tag steps
  local lib=multimot_track_b200/liborbx.so pad=0
  [ $2 = prev ] && lib=$prev
  [ $2 = new4 ] && lib=multimot_track_b200/liborbx_dbg.so && pad=10
  ORBX_FAST_PAD_KB=$pad ORBX_LIBRARY=$lib timeout 150 python bench.py --workload $1 --steps $4 --warmup 20 --no-cpu-baseline --no-extras \
      > gpurun_out/ab_$1_$2_$3.json 2> gpurun_out/ab_$1_$2_$3.err
  python - <<P >> $S
import json
try:
    d = json.loads(open("gpurun_out/ab_$1_$2_$3.json").read().strip().splitlines()[-1])
    print("$1 $2 $3 value %.0f e2e %.0f ms %.4f fast %.4f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["stage_ms"]["fast"], d.get("parity_checked")))
except Exception as e:
    print("$1 $2 $3 failed", e)
P
}
one k1 new 1 1000; one k1 prev 1 1000; one k1 new4 1 1000; one k1 new 2 1000; one k1 prev 2 1000
one k2 new 1 200
cat $S
