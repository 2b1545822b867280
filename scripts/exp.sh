run() { echo "== $1"; shift; env "$@" python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']), round(d['e2e']['value']), d['roofline']['stage_ms'])"; }
run split A=1
run nosplit ORBX_OCT_NOSPLIT=1
run split80 ORBX_OCT_BUDGET_KB=80
run split_again A=1
