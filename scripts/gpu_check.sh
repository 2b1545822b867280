#!/bin/bash
# scripts/gpu_check.sh [tag] -- on the GPU box: parity tests, then the default bench; outputs under gpurun_out/
tag=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/pytest_$tag.log
python bench.py --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo bench_exit=$?
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json")); print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["stage_ms"])
except Exception as e: print("bench parse failed", e)
PY
