"""The JSON lines bench.py printed on the B200 (committed under profiles/) carry every key the measurement contract names.
Guards the contract against regressions without needing a GPU; the live run is exercised by the driver."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip("no committed bench line " + name)
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith("{"):
            return json.loads(ln)
    raise AssertionError("no JSON line in " + name)


BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e")


@pytest.mark.parametrize("name", ["r1_bench_k1.json", "r1_bench_k2.json", "r1_bench_k4.json", "r1_bench_k1_8gpu.json",
                                  "r2_bench_k1.json", "r2_bench_k1_20steps.json", "r2_bench_k2.json", "r2_bench_k4.json", "r2_bench_k1_8gpu.json",
                                  "r2_bench_k1_2gpu.json", "r2_bench_k1_4gpu_20steps.json", "r2_bench_k4_2gpu.json", "r2_bench_k4_4gpu.json", "r2_bench_k4_8gpu.json",
                                  "r2b_bench_k1.json", "r2b_bench_k1_20steps.json", "r2b_bench_k2.json", "r2b_bench_k4.json"])
def test_our_arm_line(name):
    d = _line(name)
    for k in BASE_KEYS + ("clocks", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["metric"] == "orb_extract_describe_frames_per_s" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["value"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if d["n_gpus"] == 1 and "cpu_baseline" in d and d["cpu_baseline"]:
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, k
        assert c["kind"] in ("reference", "port")
    if name.startswith("r2"):
        assert d["parity_checked"] is True and d["parity"]["frames_with_keypoint_mismatch"] == 0 and d["parity"]["descriptor_bits_differing"] == 0
        if d["n_gpus"] == 1 and d.get("cpu_baseline"):
            legs = d["cpu_baseline"]["legs"]                       # both CPU legs, the faster one named
            assert {"reference", "cv2"} <= set(legs) and d["cpu_baseline"]["value"] == max(v["value"] for v in legs.values() if "value" in v)
        if d.get("matcher") and "error" not in d["matcher"]:
            m = d["matcher"]
            assert m["equals_cpu_scan"] is True and m["accepted_th_high"] >= m["accepted_th_low"] > 0 and 0 < m["frac_of_popc_peak"] < 1
            assert m["cpu_baseline"]["ms_per_call"] > m["ms_per_call"]


@pytest.mark.parametrize("name", ["r1_bench_reference_arm.json", "r2_bench_reference_arm.json"])
def test_reference_arm_line(name):
    d = _line(name)
    for k in BASE_KEYS + ("impl", "cpu_baseline"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "orb_extract_describe_frames_per_s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]


def test_algorithmic_bytes_formula_matches_survey():
    """B_alg = 5P + P0 - P_last + 16C + 96K + min(749K, P) + min(512K, P); K1 exact: 10.59 MB per frame."""
    import sys
    sys.path.insert(0, ROOT)
    from bench import algorithmic_bytes
    k1 = [(1242, 375), (1035, 312), (862, 260), (719, 217), (599, 181), (499, 151), (416, 126), (347, 105)]
    C, K = 18410, 2008
    P = sum(w * h for w, h in k1)
    assert P == 1441432
    st = algorithmic_bytes(k1, C, K)
    expect = 5 * P + k1[0][0] * k1[0][1] - k1[-1][0] * k1[-1][1] + 16 * C + 96 * K + min(749 * K, P) + min(512 * K, P)
    assert st["total"] == expect and abs(st["total"] / 1e6 - 10.59) < 0.01
    assert st["pyramid"] + st["fast"] + st["octree"] + st["orient_desc"] + st["blur"] == st["total"]
