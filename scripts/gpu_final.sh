#!/bin/bash
# scripts/gpu_final.sh [tag] -- short validation + evidence run after a kernel change (one GPU, about four minutes):
# GPU tests, smoke, the default bench line, the driver-style 20-step line, K2 / K4 lines, one `ncu --set full` capture of k_fast_fused.
tag=${1:-r2b}
O=gpurun_out; mkdir -p $O
timeout 120 python -m pytest tests -m gpu -x -q > $O/final_${tag}_tests.log 2>&1; echo tests_exit=$?; tail -1 $O/final_${tag}_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 200 python bench.py > $O/bench_${tag}_k1.json 2> $O/bench_${tag}_k1.err; echo bench_k1_exit=$?
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${tag}_k1_20steps.json 2>/dev/null; echo bench_k1_20_exit=$?
timeout 80 python bench.py --workload k2 --steps 200 --no-cpu-baseline --no-extras > $O/bench_${tag}_k2.json 2> $O/bench_${tag}_k2.err; echo bench_k2_exit=$?
timeout 80 python bench.py --workload k4 --steps 300 --no-cpu-baseline --no-extras > $O/bench_${tag}_k4.json 2> $O/bench_${tag}_k4.err; echo bench_k4_exit=$?
timeout 120 ncu --set full --import-source on --clock-control none -k regex:k_fast_fused -s 4 -c 1 -f -o $O/prof_${tag}_fast \
    python scripts/run_workload.py k1 4 > $O/ncu_full_${tag}_fast.log 2>&1; echo ncu_fast_exit=$?
ncu -i $O/prof_${tag}_fast.ncu-rep --page raw --csv > $O/ncu_raw_${tag}_fast.csv 2>/dev/null
echo "-- hottest source lines, k_fast_fused" > $O/ncu_hot_${tag}_fast.txt
python scripts/ncu_source_hot.py $O/prof_${tag}_fast.ncu-rep k_fast_fused 16 >> $O/ncu_hot_${tag}_fast.txt 2>/dev/null
rm -f $O/prof_${tag}_fast.ncu-rep
python - <<P
import json
for n in ("k1", "k1_20steps", "k2", "k4"):
    try:
        d = json.loads(open("$O/bench_${tag}_%s.json" % n).read().strip().splitlines()[-1])
        print(n, "value %.0f e2e %.0f ms %.4f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("parity_checked")))
    except Exception as e:
        print(n, "failed", e)
P
