#!/bin/bash
# The persistent TMA-staged FAST variant against the default kernel in the pipelined loop (K1, then K2), each with bench.py's parity self-check.
mkdir -p gpurun_out; S=gpurun_out/tma_check.txt; : > $S
one() {
  timeout 40 python bench.py --workload $1 --steps $2 --warmup 10 --no-cpu-baseline --no-extras $3 > gpurun_out/tma_$1_$4.json 2>/dev/null
  python - <<P >> $S
import json
try:
    d = json.loads(open("gpurun_out/tma_$1_$4.json").read().strip().splitlines()[-1])
    print("$1 $4 value %.0f ms %.4f fast-alone %.4f parity %s" % (d["value"], d["ms_per_step"], d["roofline"]["stage_ms"]["fast"], d.get("parity_checked")))
except Exception as e:
    print("$1 $4 failed", e)
P
}
one k1 1000 "" default; one k1 1000 --fast-tma tma; one k2 100 --fast-tma tma; one k2 100 "" default
cat $S
