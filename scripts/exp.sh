run() { echo "== $1"; shift; env "$@" python bench.py --no-cpu-baseline --handles 6 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']), round(d['e2e']['value']), d['roofline']['stage_ms'])"; }
run v0 ORBX_NO_TMA_FAST=1 ORBX_FF_VARIANT=0
run v1_5ctas ORBX_NO_TMA_FAST=1 ORBX_FF_VARIANT=1
run v2_unr12 ORBX_NO_TMA_FAST=1 ORBX_FF_VARIANT=2
run v3_both ORBX_NO_TMA_FAST=1 ORBX_FF_VARIANT=3
run tma A=1
