"""Pins the oracle port (oracle/orb_oracle.c) to the reference:
  * golden outputs of the reference's own ORBextractor.cc (canonical tie-break build)
    on the two committed KITTI sample frames, both parameter sets (tests/golden/);
  * live, when oracle/_ref/*.so is present (built from /root/reference here; shipped
    to the GPU box), on synthetic frames too, incl. DescriptorDistance."""
import numpy as np
import pytest

from conftest import crc32, desc_bit_mismatch, kps_equal_exact


@pytest.mark.parametrize("frame", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("nfeat", [2000, 4000])
def test_port_matches_reference_golden(oracle_mod, golden_kitti, kitti_frames, frame, nfeat):
    tag = "f%d_n%d" % (frame, nfeat)
    o = oracle_mod.Oracle(nfeat, 1.2, 8, 20, 7)
    kps, desc = o(kitti_frames[frame])
    ref_k, ref_d = golden_kitti["kps_" + tag], golden_kitti["desc_" + tag]
    assert kps_equal_exact(kps, ref_k)
    assert np.array_equal(kps["angle"].view(np.uint32), ref_k["angle"].view(np.uint32))      # same fastAtan2 polynomial
    bad, total = desc_bit_mismatch(desc, ref_d)
    assert bad == 0, "descriptor bits differing from the reference: %d of %d" % (bad, total)
    assert [crc32(o.level_image(l)) for l in range(8)] == golden_kitti["pyr_crc_" + tag].tolist()
    assert [crc32(o.level_padded(l)) for l in range(8)] == golden_kitti["pyr_border_crc_" + tag].tolist()
    assert [len(o.level_candidates(l)) for l in range(8)] == golden_kitti["ncand_" + tag].tolist()


def test_survey_anchors(golden_kitti):
    """Appendix B of SURVEY.md (measured with cv2 during the survey) agrees with the reference-derived goldens."""
    assert "%08x" % int(golden_kitti["gray_crc_0"]) == "afa276f3"
    assert ["%08x" % v for v in golden_kitti["pyr_crc_f0_n2000"]] == \
        ["afa276f3", "69c746e0", "6c793fe7", "a1c81c73", "a30f55d4", "d2064ff0", "0e9b1e9b", "bfb93a19"]
    assert ["%08x" % v for v in golden_kitti["blur_crc_f0_n2000"]] == \
        ["fe8ec1e9", "e84cd204", "d6186819", "dad34ae4", "7cf9a9ec", "bbad43fe", "30bf3a11", "98e84fec"]
    assert golden_kitti["ncand_f0_n2000"].tolist() == [7507, 5158, 3539, 2461, 1682, 1148, 794, 533]
    assert golden_kitti["mincells_f0_n2000"].tolist() == [109, 69, 31, 18, 6, 0, 0, 0]
    assert np.bincount(golden_kitti["kps_f0_n2000"]["octave"]).tolist() == [436, 364, 303, 251, 211, 176, 145, 122]


def test_constructor_tables(oracle_mod, golden_kitti):
    t = oracle_mod.Oracle(2000, 1.2, 8, 20, 7).tables()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
        assert np.array_equal(t[k].view(np.uint32), golden_kitti["tables2000_" + k].view(np.uint32)), k
    assert t["nfeat"].tolist() == [434, 362, 302, 251, 209, 175, 145, 122]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    t12 = oracle_mod.Oracle(10000, 1.2, 12, 20, 7).tables()
    assert np.array_equal(t12["nfeat"], golden_kitti["tables10000x12_nfeat"])
    assert np.array_equal(t12["scale"].view(np.uint32), golden_kitti["tables10000x12_scale"].view(np.uint32))


def test_synthetic_golden(oracle_mod, golden_synth):
    pytest.importorskip("cv2")
    from multimot_track_b200.synth import value_noise_frame
    for (h, w) in ((375, 1242), (1080, 1920)):
        assert crc32(value_noise_frame(0, h, w)) == int(golden_synth["crc_%dx%d_seed0" % (w, h)])
    assert "%08x" % int(golden_synth["crc_1242x375_seed0"]) == "dc51c799"
    o = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)
    for seed in (0, 1):
        kps, desc = o(value_noise_frame(seed, 375, 1242))
        assert kps_equal_exact(kps, golden_synth["kps_1242x375_seed%d" % seed])
        assert np.array_equal(desc, golden_synth["desc_1242x375_seed%d" % seed])


def test_port_matches_live_reference(oracle_mod):
    if not oracle_mod.RefExtractor.available("canon"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(11)
    frames = [rng.integers(0, 256, (240, 320), dtype=np.uint8)]
    try:
        from multimot_track_b200.synth import value_noise_frame
        frames += [value_noise_frame(21, 300, 500), value_noise_frame(22, 480, 640)]
    except ImportError:
        pass
    for img in frames:
        for nfeat, L in ((1000, 5), (300, 3)):
            ref = oracle_mod.RefExtractor(nfeat, 1.2, L, 20, 7, "canon")
            o = oracle_mod.Oracle(nfeat, 1.2, L, 20, 7)
            rk, rd = ref(img)
            ok, od = o(img)
            assert kps_equal_exact(ok, rk) and np.array_equal(ok["angle"].view(np.uint32), rk["angle"].view(np.uint32))
            assert np.array_equal(od, rd)
            for l in range(L):
                assert np.array_equal(o.level_image(l), ref.pyramid_level(l))
                assert np.array_equal(o.level_padded(l), ref.pyramid_level(l, True))


@pytest.mark.parametrize("shape,params", [((241, 323), (400, 1.7, 3, 20, 7)), ((613, 401), (800, 1.85, 3, 20, 7)),
                                          ((377, 1241), (1200, 2.0, 3, 20, 7)), ((301, 1023), (900, 1.1, 10, 30, 2)),
                                          ((230, 231), (300, 1.2, 8, 20, 7)), ((300, 700), (500, 1.5, 4, 25, 10))])
def test_port_matches_live_reference_scale_factors(oracle_mod, shape, params):
    """The shapes / scale factors / thresholds of the GPU parity tests (tests/test_gpu_parity.py::test_scale_factors_and_odd_sizes,
    test_other_shapes_and_parameters): the port equals the reference's own ORBextractor.cc there too, on value noise and on dense
    uniform noise (keypoints incl. angle bits, descriptors, every pyramid level with and without the border)."""
    if not oracle_mod.RefExtractor.available("canon"):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    from multimot_track_b200.synth import uniform_noise_frame, value_noise_frame
    ref = oracle_mod.RefExtractor(*params, "canon")
    o = oracle_mod.Oracle(*params)
    for img in (value_noise_frame(41, *shape), uniform_noise_frame(6, *shape)):
        rk, rd = ref(img)
        ok, od = o(img)
        assert len(ok) > 50 and kps_equal_exact(ok, rk) and np.array_equal(ok["angle"].view(np.uint32), rk["angle"].view(np.uint32))
        assert np.array_equal(od, rd)
        for l in range(params[2]):
            assert np.array_equal(o.level_image(l), ref.pyramid_level(l))
            assert np.array_equal(o.level_padded(l), ref.pyramid_level(l, True))


def test_hamming_and_matcher_rule(oracle_mod):
    rng = np.random.default_rng(3)
    A = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (55, 32), dtype=np.uint8)
    B[7] = A[3]; B[20] = A[3]                        # duplicate minimum: second-best equals best
    d = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(axis=2)
    for i in range(5):
        for j in range(5):
            assert oracle_mod.Oracle.hamming256(A[i], B[j]) == d[i, j]
            if oracle_mod.RefExtractor.available("canon"):
                assert oracle_mod.RefExtractor.descriptor_distance(A[i], B[j]) == d[i, j]
    idx, d1, d2, acc = oracle_mod.Oracle.match(A, B, 100, 0.9)
    assert np.array_equal(idx, d.argmin(axis=1))     # argmin returns the lowest index on ties
    assert np.array_equal(d1, d.min(axis=1))
    assert np.array_equal(d2, np.sort(d, axis=1)[:, 1])
    assert idx[3] == 7 and d1[3] == 0 and d2[3] == 0 and not acc[3]      # 0 < 0.9*0 is false
    exp = (d1 <= 100) & (d1.astype(np.float32) < np.float32(0.9) * d2.astype(np.float32))
    assert np.array_equal(acc, exp)
    i2, a1, a2, ac2 = oracle_mod.Oracle.match(A, B, 100, 0.9, threads=3)
    assert np.array_equal(i2, idx) and np.array_equal(a1, d1) and np.array_equal(a2, d2) and np.array_equal(ac2, acc)
    e_idx, e_d1, e_d2, e_acc = oracle_mod.Oracle.match(A, np.zeros((0, 32), np.uint8), 100, 0.9)
    assert (e_idx == -1).all() and (e_d1 == 256).all() and not e_acc.any()


def _rotation_cases():
    rng = np.random.default_rng(17)
    cases = []
    for n, nB, mode in ((400, 380, "spread"), (900, 900, "peaked"), (64, 50, "ties"), (7, 9, "tiny"), (0, 5, "empty")):
        idx = rng.integers(0, max(nB, 1), n).astype(np.int32)
        acc = rng.random(n) < 0.7
        b = (rng.random(nB) * 360).astype(np.float32)
        if mode == "peaked":                      # most matches rotate by ~12 degrees, a few outliers
            a = (b[idx] + np.float32(12) + rng.normal(0, 3, n).astype(np.float32)) % np.float32(360)
            out = rng.random(n) < 0.15
            a[out] = (rng.random(int(out.sum())) * 360).astype(np.float32)
        elif mode == "ties":                      # equal bin counts: the first bin must win
            a = (b[idx] + np.float32(30) * (np.arange(n) % 4).astype(np.float32) + np.float32(1)) % np.float32(360)
            acc[:] = True
        else:
            a = (rng.random(n) * 360).astype(np.float32)
        a = a.astype(np.float32)
        a[::11] = np.float32(0); b[::13] = np.float32(359.99997)      # wrap-around and exact-boundary angles
        if n:
            a[-1] = b[idx[-1]]                    # rot == 0
        cases.append((idx, acc, a, b))
    return cases


def test_rotation_filter_port_vs_reference(oracle_mod):
    """mbCheckOrientation: the port against a numpy restatement, and against the reference's own compiled
    ComputeThreeMaxima (src/ORBmatcher.cc:2233-2274) when /root/reference is here."""
    for idx, acc, a, b in _rotation_cases():
        k_acc, hist, top3 = oracle_mod.Oracle.rotation_filter(idx, acc, a, b)
        # numpy restatement of :610-620 (float32 arithmetic, round half away from zero)
        rot = (a - b[idx] if len(idx) else np.zeros(0, np.float32)).astype(np.float32)
        rot = np.where(rot < 0, (rot + np.float32(360)).astype(np.float32), rot)
        t = (rot * np.float32(1.0 / 30)).astype(np.float32)
        bins = np.floor(t.astype(np.float64) + 0.5).astype(np.int64)
        bins[bins == 30] = 0
        exp_hist = np.bincount(bins[acc], minlength=30)[:30] if len(idx) else np.zeros(30, np.int64)
        assert np.array_equal(hist, exp_hist)
        assert hist[13:].sum() == 0               # factor = 1/30: the reference only ever fills bins 0..12
        keep = np.isin(bins, top3[top3 >= 0]) & acc if len(idx) else acc
        assert np.array_equal(k_acc, keep)
        if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_rotation_filter"):
            r_acc, r_hist, r_top3 = oracle_mod.RefExtractor.rotation_filter(idx, acc, a, b)
            assert np.array_equal(r_acc, k_acc) and np.array_equal(r_hist, hist) and np.array_equal(r_top3, top3)


def _stereo_inputs(oracle_mod, seed, h, w, params):
    from multimot_track_b200.synth import stereo_pair
    L, R = stereo_pair(seed, h, w)
    oL, oR = oracle_mod.Oracle(*params), oracle_mod.Oracle(*params)
    kL, dL = oL(L); kR, dR = oR(R)
    t = oL.tables()
    pyrL = [oL.level_image(l) for l in range(params[2])]; pyrR = [oR.level_image(l) for l in range(params[2])]
    return L, R, kL, dL, kR, dR, t, pyrL, pyrR


@pytest.mark.parametrize("seed,shape,params", [(3, (375, 1242), (2000, 1.2, 8, 20, 7)), (5, (240, 400), (600, 1.2, 4, 20, 7))])
def test_stereo_port_vs_reference(oracle_mod, seed, shape, params):
    """Frame::ComputeStereoMatches: the C port against the reference's own compiled function body (src/Frame.cc:849-1038
    excerpted unmodified, oracle/Makefile) on a synthetic rectified pair; bit-identical mvuRight / mvDepth / vDescIndex."""
    _, _, kL, dL, kR, dR, t, pyrL, pyrR = _stereo_inputs(oracle_mod, seed, shape[0], shape[1], params)
    a = oracle_mod.Oracle.stereo_matches(kL, dL, kR, dR, t["scale"], t["inv_scale"], pyrL, pyrR, 386.1448)
    assert a["kept"] > 30                                  # the pair really produces stereo matches
    m = a["u_right"] >= 0
    disp = kL["x"][m] - a["u_right"][m]
    assert (disp > 0).all() and (disp < 200).all() and np.allclose(a["depth"][m], np.float32(386.1448) / disp, rtol=1e-6)
    assert (a["desc_index"] != 0).all()                    # the `bestIdxR != 0` quirk (:948)
    assert ((a["sad"] >= 0) >= m).all() and (a["u_right"][a["best_dist"] >= 75] < 0).all()
    if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_stereo_matches"):
        b = oracle_mod.RefExtractor.stereo_matches(kL, dL, kR, dR, t["scale"], t["inv_scale"], pyrL, pyrR, 386.1448)
        assert b["kept"] == a["kept"]
        assert np.array_equal(a["u_right"].view(np.uint32), b["u_right"].view(np.uint32))
        assert np.array_equal(a["depth"].view(np.uint32), b["depth"].view(np.uint32))
        assert np.array_equal(a["desc_index"], b["desc_index"])


def _distinctive_cases(seed=23):
    """Map points with 0..70 observations: noisy copies of a base descriptor, plus ties and identical rows."""
    rng = np.random.default_rng(seed)
    rows, offsets = [], [0]
    for n in (1, 2, 3, 4, 5, 8, 13, 0, 33, 64, 70, 2, 6):
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        d = np.repeat(base[None, :], n, 0).copy()
        for i in range(n):
            for f in rng.integers(0, 256, rng.integers(0, 14)):
                d[i, f % 32] ^= np.uint8(1 << (f % 8))
        if n == 6:
            d[3] = d[1]; d[5] = d[0]                      # exact duplicates: equal medians, the first row must win
        rows.append(d); offsets.append(offsets[-1] + n)
    return np.concatenate(rows), np.array(offsets, np.int32)


def test_distinctive_descriptors_port_vs_reference(oracle_mod):
    """MapPoint::ComputeDistinctiveDescriptors: the port against numpy and against the reference's own loop
    (src/MapPoint.cc:272-301 excerpted unmodified, oracle/Makefile)."""
    desc, off = _distinctive_cases()
    best, med = oracle_mod.Oracle.distinctive_descriptors(desc, off)
    for p in range(len(off) - 1):
        d = desc[off[p]:off[p + 1]]
        n = len(d)
        if n == 0:
            assert best[p] == -1
            continue
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(axis=2)
        medians = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert best[p] == int(np.argmin(medians)) and med[p] == medians.min()
        if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_distinctive_descriptor"):
            assert oracle_mod.RefExtractor.distinctive_descriptor(d) == (best[p], med[p])


def _voc_cases(tmp_path, seeds):
    from multimot_track_b200.synth import write_synthetic_vocabulary
    cases = []
    for i, (k, L, sc, wt) in enumerate([(10, 3, 0, 0), (9, 3, 1, 1), (6, 4, 5, 0), (10, 2, 0, 2), (4, 5, 2, 3)]):
        path = str(tmp_path / ("voc%d.txt" % i))
        write_synthetic_vocabulary(path, k, L, seed=50 + i, scoring=sc, weighting=wt, seeds=seeds)
        cases.append((path, k, L))
    return cases


def test_vocabulary_port_vs_dbow2(oracle_mod, tmp_path):
    """Frame::ComputeBoW's vocabulary (DBoW2): the C port against the reference's vendored DBoW2 compiled unmodified
    (oracle/_ref/liborbref_bow.so): word / node / weight per descriptor, BowVector values as bit patterns, FeatureVector."""
    from multimot_track_b200.synth import value_noise_frame
    _, d = oracle_mod.Oracle(800, 1.2, 6, 20, 7)(value_noise_frame(2, 300, 600))
    for path, k, L in _voc_cases(tmp_path, d):
        a = oracle_mod.OracleVocabulary(path)
        assert (a.k, a.levels, a.nodes) == (k, L, (k ** (L + 1) - 1) // (k - 1)) and a.words == k ** L
        word, node, weight = a.transform_each(d, 4)
        assert (word >= 0).all() and (word < a.words).all() and (weight >= 0).all()
        bi, bv, fn, fo, ff = a.transform(d, 2)
        assert (np.diff(bi) > 0).all() and (np.diff(fn) > 0).all() and sorted(ff.tolist()) == sorted(np.nonzero(a.transform_each(d, 2)[2] > 0)[0].tolist())
        if not oracle_mod.RefVocabulary.available():
            continue
        b = oracle_mod.RefVocabulary(path)
        assert (b.k, b.levels, b.nodes, b.words) == (a.k, a.levels, a.nodes, a.words)
        for lu in (4, 2, 1, 0):
            for x, y in zip(a.transform_each(d, lu), b.transform_each(d, lu)):
                assert np.array_equal(x, y)
            ta, tb = a.transform(d, lu), b.transform(d, lu)
            assert np.array_equal(ta[1].view(np.uint64), tb[1].view(np.uint64))
            for x, y in zip(ta, tb):
                assert np.array_equal(x, y)


def _projection_cases(oracle_mod):
    from multimot_track_b200.synth import projection_case, value_noise_frame
    o = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)
    k, d = o(value_noise_frame(0, 375, 1242))
    _, d2 = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)(value_noise_frame(1, 375, 1242))
    sc = o.tables()["scale"]
    # (seed, th, mono, forward motion): lateral / mono / forward (bForward) / backward (bBackward) / wide window
    return [projection_case(s, k, d, sc, th, mono, fw, d2[:700]) for s, th, mono, fw in
            ((1, 15.0, False, 0.0), (2, 7.0, True, 0.0), (3, 15.0, False, 1.2), (4, 30.0, False, -1.2), (5, 45.0, False, 0.0))]


def _initialization_cases(oracle_mod):
    from multimot_track_b200.synth import initialization_case, value_noise_frame
    k, d = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)(value_noise_frame(0, 375, 1242))
    _, d2 = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)(value_noise_frame(1, 375, 1242))
    # (seed, windowSize, nnratio, shift): Tracking's call (window 100, ratio 0.9), a tight window, a strict ratio, off-image windows
    return [initialization_case(s, k, d, w, r, sh, d2[:600]) for s, w, r, sh in
            ((1, 100, 0.9, (6.0, -3.0)), (2, 10, 0.9, (3.0, 2.0)), (3, 100, 0.6, (-20.0, 9.0)), (4, 40, 0.9, (700.0, 0.0)))]


def test_search_for_initialization_port_vs_reference(oracle_mod):
    """ORBmatcher::SearchForInitialization: the C port against the reference's own function (src/ORBmatcher.cc:780-895 excerpted
    unmodified, with Frame::GetFeaturesInArea): vnMatches12, vbPrevMatched after the call, nmatches; steals do happen."""
    total = 0
    for case in _initialization_cases(oracle_mod):
        m12, prev, n = oracle_mod.search_for_initialization_port(case)
        assert n == (m12 >= 0).sum()
        taken = m12[m12 >= 0]
        assert len(np.unique(taken)) == len(taken) and (case["oct1"][m12 >= 0] == 0).all() and (case["oct2"][taken] == 0).all()
        assert np.array_equal(prev[m12 >= 0], case["xy2"][taken]) and np.array_equal(prev[m12 < 0], case["prev_xy"][m12 < 0])
        total += n
        no_ori = dict(case, check_orientation=False)
        assert oracle_mod.search_for_initialization_port(no_ori)[2] >= n
        if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_search_for_initialization"):
            for c in (case, no_ori):
                a, b = oracle_mod.search_for_initialization_port(c), oracle_mod.search_for_initialization_ref(c)
                assert a[2] == b[2] and np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert total > 300


def _local_points_cases(oracle_mod):
    from multimot_track_b200.synth import local_points_case, value_noise_frame
    o = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)
    k, d = o(value_noise_frame(0, 375, 1242))
    _, d2 = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)(value_noise_frame(1, 375, 1242))
    sc = o.tables()["scale"]
    # (seed, th, nnratio): Tracking's th for RGB-D (3), monocular (1: bFactor off), after relocalisation (5); a strict ratio
    return [local_points_case(s, k, d, sc, th, r, d2[:700]) for s, th, r in ((1, 3.0, 0.8), (2, 1.0, 0.8), (3, 5.0, 0.8), (4, 3.0, 0.6))]


def test_search_local_points_port_vs_reference(oracle_mod):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th): the C port against the reference's own function (src/ORBmatcher.cc:418-511
    excerpted unmodified, with Frame::GetFeaturesInArea): the map point assigned to every feature and nmatches."""
    for case in _local_points_cases(oracle_mod):
        a, na = oracle_mod.search_local_points_port(case)
        assert na > 800 and na >= (a >= 0).sum() > 800
        assert case["valid"][a[a >= 0]].all() and not (case["feat_obs"][a >= 0] > 0).any()        # valid points only, never onto an observed one
        if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_search_local_points"):
            b, nb = oracle_mod.search_local_points_ref(case)
            assert nb == na and np.array_equal(a, b)


def _bow_match_cases(oracle_mod, node_fn=None):
    from multimot_track_b200.synth import bow_match_case, value_noise_frame
    k, d = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)(value_noise_frame(0, 375, 1242))
    _, d2 = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)(value_noise_frame(1, 375, 1242))
    kw = {} if node_fn is None else {"node_fn": node_fn}
    # (seed, nnratio): TrackReferenceKeyFrame (0.7), Relocalization (0.75), a loose and a strict ratio
    return [bow_match_case(s, k, d, r, d2[:600], **kw) for s, r in ((1, 0.7), (2, 0.75), (3, 0.9), (4, 0.6))]


def test_search_by_bow_port_vs_reference(oracle_mod):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...): the C port against the reference's own function (src/ORBmatcher.cc:532-663
    excerpted unmodified): the key-frame feature matched to every frame feature and nmatches, with and without the rotation check."""
    for case in _bow_match_cases(oracle_mod):
        for c in (case, dict(case, check_orientation=False)):
            a, na = oracle_mod.search_by_bow_port(c)
            assert na == (a >= 0).sum() > 500 and case["kf_valid"][a[a >= 0]].all()
            if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_search_by_bow"):
                b, nb = oracle_mod.search_by_bow_ref(c)
                assert nb == na and np.array_equal(a, b)


def test_search_by_bow_keyframes_port_vs_reference(oracle_mod):
    """ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12): the C port against the reference's own function
    (src/ORBmatcher.cc:897-1030 excerpted unmodified)."""
    for case in _bow_match_cases(oracle_mod):
        for c in (case, dict(case, check_orientation=False)):
            a, na = oracle_mod.search_by_bow_kf_port(c)
            taken = a[a >= 0]
            assert na == len(taken) > 400 and len(np.unique(taken)) == len(taken) and case["f_valid"][taken].all() and case["kf_valid"][a >= 0].all()
            if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_search_by_bow_kf"):
                b, nb = oracle_mod.search_by_bow_kf_ref(c)
                assert nb == na and np.array_equal(a, b)


def test_minicv_float_gemm_against_live_cv2(oracle_mod):
    """The projection matcher's `Rcw*x3Dw+tcw` and `-Rcw.t()*tcw` (src/ORBmatcher.cc:1968-1976, 1990-1991) go through cv::gemm;
    the port's arithmetic (float accumulation for A*B+C, double for the transposed product) is pinned to cv2 4.13 here."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    for _ in range(3000):
        R = rng.normal(size=(3, 3)).astype(np.float32); x = (rng.normal(size=(3, 1)) * 10).astype(np.float32); t = rng.normal(size=(3, 1)).astype(np.float32)
        f = np.zeros(3, np.float32); tr = np.zeros(3, np.float32)
        for r in range(3):
            s = np.float32(0)
            for k in range(3):
                s = np.float32(s + np.float32(R[r, k] * x[k, 0]))
            f[r] = np.float32(s + t[r, 0])
            tr[r] = np.float32(-sum(float(R[k, r]) * float(t[k, 0]) for k in range(3)))
        assert np.array_equal(cv2.gemm(R, x, 1.0, t, 1.0).ravel().view(np.uint32), f.view(np.uint32))
        assert np.array_equal(cv2.gemm(R, t, -1.0, None, 0.0, flags=cv2.GEMM_1_T).ravel().view(np.uint32), tr.view(np.uint32))


def test_search_by_projection_port_vs_reference(oracle_mod):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, ...): the C port against the reference's own function and Frame grid
    (src/ORBmatcher.cc:1958-2102, src/Frame.cc:601-616, 710-776 excerpted unmodified): the final mvpMapPoints assignment and nmatches."""
    for case in _projection_cases(oracle_mod):
        a, na = oracle_mod.search_by_projection_port(case)
        assert na > 1000 and (a >= 0).sum() > 1000
        taken = a[a >= 0]
        assert len(np.unique(taken)) == len(taken) and case["valid"][taken].all()          # one feature per map point, valid points only
        if oracle_mod.RefExtractor.available("canon") and hasattr(oracle_mod.RefExtractor.lib("canon"), "orbref_search_by_projection"):
            b, nb = oracle_mod.search_by_projection_ref(case)
            assert nb == na and np.array_equal(a, b)


def test_cv2_restatement_cross_check(oracle_mod, kitti_frames, golden_kitti):
    """the independent Python restatement on the real cv2 4.13 primitives (oracle/cv2_restatement.py; also the
    second CPU baseline of bench.py) against the C port and the reference-derived goldens: every stage and all outputs identical."""
    cv2 = pytest.importorskip("cv2")
    from oracle.cv2_restatement import Cv2Extractor
    from multimot_track_b200.synth import value_noise_frame
    for img, params, gold in ((kitti_frames[2], (4000, 1.2, 8, 20, 7), "f2_n4000"), (value_noise_frame(3, 375, 1242), (2000, 1.2, 8, 20, 7), None),
                              (value_noise_frame(5, 300, 500), (700, 1.3, 5, 15, 5), None)):
        e = Cv2Extractor(*params)
        k, d = e(img)
        o = oracle_mod.Oracle(*params)
        ok, od = o(img)
        for l in range(params[2]):
            assert np.array_equal(e.pyramid[l], o.level_image(l)) and np.array_equal(e.padded[l], o.level_padded(l)), l
            assert np.array_equal(e.candidates[l], o.level_candidates(l)), l
            if e.blurred[l] is not None:
                assert np.array_equal(e.blurred[l], o.level_blurred(l)), l
        assert kps_equal_exact(k, ok) and np.array_equal(k["angle"].view(np.uint32), ok["angle"].view(np.uint32))
        assert desc_bit_mismatch(d, od)[0] == 0
        if gold:
            assert kps_equal_exact(k, golden_kitti["kps_" + gold]) and np.array_equal(d, golden_kitti["desc_" + gold])
