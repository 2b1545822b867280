"""Host logic of the windowed / BoW-guided matchers without a GPU: multimot_track_b200/csrc/host_match.cpp (the sequential replay of
the reference's claim bookkeeping, compiled on its own with g++) is fed candidate lists built HERE in numpy the way the kernels
k_window_candidates / k_local_candidates / k_bow_pair_distances define them (float32 grid arithmetic of Frame::GetFeaturesInArea /
PosInGrid, src/Frame.cc:710-776; packed keys cell << 32 | index << 16 | distance) and must give the oracle's matches.  The GPU
tests check the same functions end to end through the C ABI; this one keeps the host half green on CPU-only boxes."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32 = np.float32
_VP, _I, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostmatch") / "libhostmatch.so")
    src = os.path.join(ROOT, "multimot_track_b200", "csrc", "host_match.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", out, src], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("g++ could not build host_match.cpp: " + r.stderr[:200])
    lib = ctypes.CDLL(out)
    names = subprocess.run(["nm", "-D", out], capture_output=True, text=True).stdout
    sym = {}
    for key in ("resolve_initialization_matches", "resolve_local_matches", "resolve_bow_matches_kf", "resolve_bow_matches", "resolve_projection_matches"):
        cand = [ln.split()[-1] for ln in names.splitlines() if " T " in ln and key in ln]
        cand.sort(key=len)                                             # resolve_bow_matches is a prefix of resolve_bow_matches_kf
        sym[key] = getattr(lib, cand[0])
        sym[key].restype = _I
    return sym


def popcount_rows(a, b):
    return np.unpackbits(np.bitwise_xor(a, b), axis=-1).sum(-1).astype(np.int64)


def roundf(v):
    """C roundf (halves away from zero) of float32 values."""
    v = v.astype(np.float64)
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int64)


def window_candidates(cam, u, v, radius, live, min_level, max_level, xy, octave, qdesc, desc, ur=None, uright=None):
    """Per query i: sorted packed keys of the features inside its window, as window_scan (kernels.cu) emits them."""
    minX, maxX, minY, maxY = (F32(c) for c in cam[6:10])
    w_inv, h_inv = F32(64.0) / (maxX - minX), F32(48.0) / (maxY - minY)
    x, y = xy[:, 0].astype(F32), xy[:, 1].astype(F32)
    px, py = roundf((x - minX) * w_inv), roundf((y - minY) * h_inv)
    lists, counts = [], []
    for i in range(len(u)):
        if not live[i]:
            counts.append(0); continue
        r = F32(radius[i])
        min_cx = max(0, int(np.floor(((F32(u[i]) - minX) - r) * w_inv))); max_cx = min(63, int(np.ceil(((F32(u[i]) - minX) + r) * w_inv)))
        min_cy = max(0, int(np.floor(((F32(v[i]) - minY) - r) * h_inv))); max_cy = min(47, int(np.ceil(((F32(v[i]) - minY) + r) * h_inv)))
        if not (min_cx < 64 and max_cx >= 0 and min_cy < 48 and max_cy >= 0):
            counts.append(0); continue
        ok = (px >= min_cx) & (px <= max_cx) & (py >= min_cy) & (py <= max_cy) & (px < 64) & (py < 48)
        if min_level[i] > 0 or max_level[i] >= 0:
            ok &= ~(octave < min_level[i])
            if max_level[i] >= 0:
                ok &= ~(octave > max_level[i])
        ok &= (np.abs(x - F32(u[i])) < r) & (np.abs(y - F32(v[i])) < r)
        if ur is not None:
            ok &= ~((uright > 0) & (np.abs(F32(ur[i]) - uright) > r))
        idx = np.nonzero(ok)[0]
        d = popcount_rows(qdesc[i][None, :], desc[idx])
        keys = ((px[idx] * 48 + py[idx]) << 32) | (idx.astype(np.int64) << 16) | d
        lists.append(np.sort(keys).astype(np.uint64)); counts.append(len(idx))
    count = np.asarray(counts, np.int32)
    offset = (np.cumsum(count) - count).astype(np.int32)
    cand = np.concatenate(lists) if lists else np.zeros(0, np.uint64)
    return np.ascontiguousarray(cand if len(cand) else np.zeros(1, np.uint64)), count, offset


def test_initialization_replay(hostlib, oracle_mod):
    from test_oracle_vs_ref import _initialization_cases
    for case in _initialization_cases(oracle_mod):
        for chk in (1, 0):
            c = dict(case, check_orientation=bool(chk))
            exp, _, nexp = oracle_mod.search_for_initialization_port(c)
            n1 = len(c["oct1"])
            live = c["oct1"] <= 0
            cand, count, offset = window_candidates(c["cam"], c["prev_xy"][:, 0], c["prev_xy"][:, 1], np.full(n1, F32(c["window"])), live,
                                                    c["oct1"], c["oct1"], c["xy2"], c["oct2"], c["desc1"], c["desc2"])
            m12 = np.zeros(n1, np.int32)
            a1, a2 = np.ascontiguousarray(c["ang1"], F32), np.ascontiguousarray(c["ang2"], F32)
            fn = hostlib["resolve_initialization_matches"]
            fn.argtypes = [_I, _I, _VP, _VP, _VP, _VP, _VP, _F, _I, _VP]
            n = fn(n1, len(c["xy2"]), cand.ctypes.data, count.ctypes.data, offset.ctypes.data, a1.ctypes.data, a2.ctypes.data, float(c["nnratio"]), chk, m12.ctypes.data)
            assert n == nexp and np.array_equal(m12, exp)


def test_local_points_replay(hostlib, oracle_mod):
    from test_oracle_vs_ref import _local_points_cases
    for c in _local_points_cases(oracle_mod):
        exp, nexp = oracle_mod.search_local_points_port(c)
        r = np.where(c["view_cos"].astype(np.float64) > 0.998, F32(2.5), F32(4.0)).astype(F32)
        if c["th"] != 1.0:
            r = (r * F32(c["th"])).astype(F32)
        radius = (r * c["scale"][c["level"]]).astype(F32)
        cand, count, offset = window_candidates(c["cam"], c["proj"][:, 0], c["proj"][:, 1], radius, c["valid"] != 0, c["level"] - 1, c["level"],
                                                c["xy"], c["octave"], c["mp_desc"], c["desc"], ur=c["proj"][:, 2], uright=c["uright"].astype(F32))
        out = np.zeros(len(c["xy"]), np.int32)
        fn = hostlib["resolve_local_matches"]
        fn.argtypes = [_I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _F, _VP]
        nobs, octv, fobs = (np.ascontiguousarray(c[k], np.int32) for k in ("nobs", "octave", "feat_obs"))
        n = fn(len(c["proj"]), len(c["xy"]), cand.ctypes.data, count.ctypes.data, offset.ctypes.data, nobs.ctypes.data, octv.ctypes.data, fobs.ctypes.data,
               float(c["nnratio"]), out.ctypes.data)
        assert n == nexp and np.array_equal(out, exp)


def _bow_entries(c, side2_all=True):
    """The merge walk of orbx_api.cu::bow_pair_distances and the pair distances of k_bow_pair_distances."""
    entries, dist, a, b, at = [], [], 0, 0, 0
    while a < len(c["kf_nodes"]) and b < len(c["f_nodes"]):
        if c["kf_nodes"][a] < c["f_nodes"][b]:
            a += 1; continue
        if c["kf_nodes"][a] > c["f_nodes"][b]:
            b += 1; continue
        lo, cnt = int(c["f_off"][b]), int(c["f_off"][b + 1] - c["f_off"][b])
        feats = c["f_feats"][lo:lo + cnt]
        for p in range(c["kf_off"][a], c["kf_off"][a + 1]):
            ik = int(c["kf_feats"][p])
            if not c["kf_valid"][ik] or cnt == 0:
                continue
            entries += [ik, lo, cnt, at]
            dist.append(popcount_rows(c["kf_desc"][ik][None, :], c["f_desc"][feats]).astype(np.uint16))
            at += cnt
        a += 1; b += 1
    return np.asarray(entries, np.int32), (np.concatenate(dist) if dist else np.zeros(1, np.uint16))


def test_bow_replay(hostlib, oracle_mod):
    from test_oracle_vs_ref import _bow_match_cases
    for case in _bow_match_cases(oracle_mod):
        entries, dist = _bow_entries(case)
        ka, fa = np.ascontiguousarray(case["kf_angle"], F32), np.ascontiguousarray(case["f_angle"], F32)
        ff, fv = np.ascontiguousarray(case["f_feats"], np.int32), np.ascontiguousarray(case["f_valid"], np.uint8)
        for chk in (1, 0):
            c = dict(case, check_orientation=bool(chk))
            exp, nexp = oracle_mod.search_by_bow_port(c)
            out = np.zeros(len(c["f_desc"]), np.int32)
            fn = hostlib["resolve_bow_matches"]
            fn.argtypes = [_I, _VP, _VP, _VP, _I, _VP, _VP, _F, _I, _VP]
            n = fn(len(entries) // 4, entries.ctypes.data, dist.ctypes.data, ff.ctypes.data, len(c["f_desc"]), ka.ctypes.data, fa.ctypes.data,
                   float(c["nnratio"]), chk, out.ctypes.data)
            assert n == nexp and np.array_equal(out, exp)
            exp, nexp = oracle_mod.search_by_bow_kf_port(c)
            m12 = np.zeros(len(c["kf_desc"]), np.int32)
            fn = hostlib["resolve_bow_matches_kf"]
            fn.argtypes = [_I, _VP, _VP, _VP, _I, _I, _VP, _VP, _VP, _F, _I, _VP]
            n = fn(len(entries) // 4, entries.ctypes.data, dist.ctypes.data, ff.ctypes.data, len(c["kf_desc"]), len(c["f_desc"]), fv.ctypes.data,
                   ka.ctypes.data, fa.ctypes.data, float(c["nnratio"]), chk, m12.ctypes.data)
            assert n == nexp and np.array_equal(m12, exp)


def test_projection_replay(hostlib, oracle_mod):
    """SearchByProjection(CurrentFrame, LastFrame): projection in cv::gemm's float order (k_project_candidates), then the host walk."""
    from test_oracle_vs_ref import _projection_cases
    for case in _projection_cases(oracle_mod):
        cam, Tc, Tl = case["cam"].astype(F32), case["Tcw_cur"].astype(F32).reshape(4, 4), case["Tcw_last"].astype(F32).reshape(4, 4)
        fx, fy, cx, cy, bf, b = cam[:6]
        # bForward / bBackward (src/ORBmatcher.cc:1968-1979): twc = -Rcw.t()*tcw (double accumulation), tlc = Rlw*twc + tlw (float)
        twc = np.array([F32(-sum(float(Tc[k, r]) * float(Tc[k, 3]) for k in range(3))) for r in range(3)], F32)
        s = F32(0)
        for k in range(3):
            s = F32(s + F32(Tl[2, k] * twc[k]))
        tlc2 = F32(s + Tl[2, 3])
        forward, backward = (tlc2 > b) and not case["mono"], (-tlc2 > b) and not case["mono"]
        W = case["world_pos"].astype(F32)
        n = len(W)
        xc3 = np.zeros((n, 3), F32)
        for r in range(3):
            acc = (Tc[r, 0] * W[:, 0]).astype(F32)
            acc = (acc + (Tc[r, 1] * W[:, 1]).astype(F32)).astype(F32)
            acc = (acc + (Tc[r, 2] * W[:, 2]).astype(F32)).astype(F32)
            xc3[:, r] = (acc + Tc[r, 3]).astype(F32)
        with np.errstate(divide="ignore"):
            invz = (1.0 / xc3[:, 2].astype(np.float64)).astype(F32)
        u = (((fx * xc3[:, 0]).astype(F32) * invz).astype(F32) + cx).astype(F32)
        v = (((fy * xc3[:, 1]).astype(F32) * invz).astype(F32) + cy).astype(F32)
        live = (case["valid"] != 0) & ~(invz < 0) & ~((u < cam[6]) | (u > cam[7])) & ~((v < cam[8]) | (v > cam[9]))
        octv = case["last_octave"]
        radius = (F32(case["th"]) * case["scale"][octv]).astype(F32)
        if forward:
            lo, hi = octv, np.full(n, -1)
        elif backward:
            lo, hi = np.zeros(n, np.int64), octv
        else:
            lo, hi = octv - 1, octv + 1
        ur = (u - (bf * invz).astype(F32)).astype(F32)
        cand, count, offset = window_candidates(cam, u, v, radius, live, lo, hi, case["cur_xy"], case["cur_octave"], case["mp_desc"], case["cur_desc"],
                                                ur=ur, uright=case["cur_uright"].astype(F32))
        la, ca = np.ascontiguousarray(case["last_angle"], F32), np.ascontiguousarray(case["cur_angle"], F32)
        nobs = np.ascontiguousarray(case["nobs"], np.int32)
        for chk in (1, 0):
            exp, nexp = oracle_mod.search_by_projection_port(dict(case, check_orientation=bool(chk)))
            out = np.zeros(len(case["cur_xy"]), np.int32)
            fn = hostlib["resolve_projection_matches"]
            fn.argtypes = [_I, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP]
            got = fn(n, len(case["cur_xy"]), cand.ctypes.data, count.ctypes.data, offset.ctypes.data, nobs.ctypes.data, la.ctypes.data, ca.ctypes.data,
                     chk, out.ctypes.data)
            assert got == nexp and np.array_equal(out, exp)
