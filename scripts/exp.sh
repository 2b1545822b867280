run() { echo "== $1"; shift; env "$@" python bench.py --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(round(d['value']), round(d['e2e']['value']), d['roofline']['stage_ms'])"; }
run base A=1
run ff4 ORBX_FF_CTAS_PER_SM=4
run ff3 ORBX_FF_CTAS_PER_SM=3
run ff2 ORBX_FF_CTAS_PER_SM=2
run ff3_bl2_od2 ORBX_FF_CTAS_PER_SM=3 ORBX_BLUR_CTAS_PER_SM=2 ORBX_OD_CTAS_PER_SM=2
run ff2_bl2_od2 ORBX_FF_CTAS_PER_SM=2 ORBX_BLUR_CTAS_PER_SM=2 ORBX_OD_CTAS_PER_SM=2
run bl2_od2 ORBX_BLUR_CTAS_PER_SM=2 ORBX_OD_CTAS_PER_SM=2
