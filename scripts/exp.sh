run() { echo "== $1"; shift; env "$@" python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']), round(d['e2e']['value']), d['roofline']['stage_ms'])"; }
run persist A=1
run iters1 ORBX_RZ_ITERS=1
run iters2 ORBX_RZ_ITERS=2
run persist_again A=1
