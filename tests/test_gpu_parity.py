"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes mirror in
multimot_track_b200), against the CPU oracle on the same inputs and against the committed
reference-derived goldens.  Bars (BASELINE.json north_star): pyramid, blur, FAST
responses, post-octree keypoints and match indices/distances bit-exact; angles within
1e-4 rad; descriptor bit mismatch rate <= 1e-4 (reported)."""
import numpy as np
import pytest

from conftest import angle_diff_rad, crc32, desc_bit_mismatch, kps_equal_exact

pytestmark = pytest.mark.gpu

ANGLE_TOL_RAD = 1e-4
DESC_MISMATCH_TOL = 1e-4


@pytest.fixture(scope="module")
def orb():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    import multimot_track_b200
    multimot_track_b200.load_library()          # raises when the CUDA extension is missing: no silent fallback
    return multimot_track_b200


def synth(seed, h, w):
    from multimot_track_b200.synth import value_noise_frame
    return value_noise_frame(seed, h, w)


def check_frame(ext, oracle, img, tag=""):
    """Full stage-by-stage comparison of one frame; returns (descriptor bits differing, total bits)."""
    kps, desc = ext(img)
    okps, odesc = oracle(img)
    L = oracle.nlevels
    for l in range(L):
        assert ext.level_size(l) == oracle.level_size(l), (tag, l)
        assert np.array_equal(ext.pyramid_level(l), oracle.level_image(l)), "%s pyramid level %d" % (tag, l)
        cand, ocand = ext.candidates(l), oracle.level_candidates(l)
        assert cand.shape == ocand.shape and np.array_equal(cand, ocand), "%s FAST candidates level %d: %d vs %d" % (tag, l, len(cand), len(ocand))
        ob = oracle.level_blurred(l)
        if ob is not None:
            assert np.array_equal(ext.blurred_level(l), ob), "%s blur level %d" % (tag, l)
    assert len(kps) == len(okps), "%s keypoint count %d vs %d" % (tag, len(kps), len(okps))
    assert kps_equal_exact(kps, okps), "%s keypoint set/order" % tag
    assert angle_diff_rad(kps["angle"], okps["angle"]).max(initial=0) <= ANGLE_TOL_RAD
    bad, total = desc_bit_mismatch(desc, odesc)
    assert bad <= DESC_MISMATCH_TOL * total, "%s descriptor mismatch %d/%d" % (tag, bad, total)
    return bad, total


def test_kitti_golden_reference(orb, golden_kitti, kitti_frames):
    """Config 1: kitti_sample frames with kitti03.yaml's (4000,1.2,8,20,7) and the benchmark's 2000,
    against the outputs of the reference's own ORBextractor.cc (tests/golden/golden_kitti.npz)."""
    bad_total = bits_total = 0
    for nfeat in (4000, 2000):
        ext = orb.ORBextractor(nfeat, 1.2, 8, 20, 7)
        for f in range(5):
            tag = "f%d_n%d" % (f, nfeat)
            kps, desc = ext(kitti_frames[f])
            ref_k, ref_d = golden_kitti["kps_" + tag], golden_kitti["desc_" + tag]
            assert [crc32(ext.pyramid_level(l)) for l in range(8)] == golden_kitti["pyr_crc_" + tag].tolist()
            assert [crc32(ext.pyramid_level(l, with_border=True)) for l in range(8)] == golden_kitti["pyr_border_crc_" + tag].tolist()
            assert [crc32(ext.blurred_level(l)) for l in range(8)] == golden_kitti["blur_crc_" + tag].tolist()
            assert [len(ext.candidates(l)) for l in range(8)] == golden_kitti["ncand_" + tag].tolist()
            assert kps_equal_exact(kps, ref_k), tag
            assert angle_diff_rad(kps["angle"], ref_k["angle"]).max() <= ANGLE_TOL_RAD
            bad, total = desc_bit_mismatch(desc, ref_d)
            bad_total += bad; bits_total += total
    print("descriptor bit mismatch vs reference: %d / %d = %.2e" % (bad_total, bits_total, bad_total / bits_total))
    assert bad_total <= DESC_MISMATCH_TOL * bits_total


def test_kitti_stages_vs_oracle(orb, oracle_mod, kitti_frames):
    ext = orb.ORBextractor(4000, 1.2, 8, 20, 7)
    o = oracle_mod.Oracle(4000, 1.2, 8, 20, 7)
    for f in range(5):
        check_frame(ext, o, kitti_frames[f], "kitti%d" % f)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_synthetic_k1_vs_oracle(orb, oracle_mod, seed):
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    o = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)
    check_frame(ext, o, synth(seed, 375, 1242), "syn%d" % seed)


def test_synthetic_golden_reference(orb, golden_synth):
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    for seed in (0, 1):
        kps, desc = ext(synth(seed, 375, 1242))
        assert kps_equal_exact(kps, golden_synth["kps_1242x375_seed%d" % seed])
        bad, total = desc_bit_mismatch(desc, golden_synth["desc_1242x375_seed%d" % seed])
        assert bad <= DESC_MISMATCH_TOL * total


def test_uniform_noise_stress(orb, oracle_mod):
    """~40k candidates on level 0: the octree keys do not fit shared memory (global scratch path)."""
    from multimot_track_b200.synth import uniform_noise_frame
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    o = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)
    check_frame(ext, o, uniform_noise_frame(5, 375, 1242), "noise")


@pytest.mark.parametrize("shape,params", [((480, 640), (1000, 1.2, 8, 20, 7)), ((240, 320), (500, 1.2, 4, 20, 7)),
                                          ((300, 700), (500, 1.5, 4, 25, 10)), ((230, 231), (300, 1.2, 8, 20, 7)),
                                          ((128, 400), (256, 1.3, 3, 15, 5)), ((500, 333), (1500, 1.2, 6, 20, 7)),
                                          ((720, 1280), (3000, 1.2, 8, 12, 4))])
def test_other_shapes_and_parameters(orb, oracle_mod, shape, params):
    ext = orb.ORBextractor(*params)
    o = oracle_mod.Oracle(*params)
    check_frame(ext, o, synth(40, *shape), "shape%s" % (shape,))


@pytest.mark.parametrize("shape,params", [((241, 323), (400, 1.7, 3, 20, 7)), ((613, 401), (800, 1.85, 3, 20, 7)),
                                          ((377, 1241), (1200, 2.0, 3, 20, 7)), ((301, 1023), (900, 1.1, 10, 30, 2))])
def test_scale_factors_and_odd_sizes(orb, oracle_mod, shape, params):
    """Odd widths / heights (partial resize, blur and FAST tiles on every side) and scale factors up to 2: 1.7 and 1.85 are the widest
    source windows the separable TMA resize takes (its byte-permute selectors span 8 source bytes per four columns), 2.0 goes to the
    plain kernel; thresholds 30 / 2 make almost every pixel a survivor candidate of the per-lane lists."""
    ext = orb.ORBextractor(*params)
    o = oracle_mod.Oracle(*params)
    check_frame(ext, o, synth(41, *shape), "scale%s" % (params[1],))
    from multimot_track_b200.synth import uniform_noise_frame
    check_frame(ext, o, uniform_noise_frame(6, *shape), "noise%s" % (shape,))


def test_edge_inputs(orb, oracle_mod):
    ext = orb.ORBextractor(1000, 1.2, 8, 20, 7)
    o = oracle_mod.Oracle(1000, 1.2, 8, 20, 7)
    # empty image: silent return (src/ORBextractor.cc:1049)
    k, d = ext(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)
    # flat image: no corners anywhere
    k, d = ext(np.full((375, 1242), 128, np.uint8))
    assert len(k) == 0
    # a single bright blob: a handful of candidates, fewer than the per-level quota
    img = np.full((375, 1242), 30, np.uint8); img[180:190, 600:612] = 220
    check_frame(ext, o, img, "blob")
    # non-CV_8UC1 input asserts like the reference (:1053)
    with pytest.raises(AssertionError):
        ext(np.zeros((375, 1242), np.float32))
    # strided view (every row padded) and a shape change on the same handle
    big = synth(8, 375, 1300)
    check_frame(ext, o, big[:, :1242], "strided")
    check_frame(ext, o, synth(9, 480, 640), "reshape")


def test_unsupported_shapes_fail_loudly(orb):
    """Shapes the reference itself cannot process (division by zero there) are refused, not guessed."""
    ext = orb.ORBextractor(1500, 1.2, 6, 20, 7)
    with pytest.raises(orb.OrbxError) as e:
        ext(synth(1, 600, 333))                      # level 5 is 134x241: round(102/209) = 0 initial octree nodes
    assert e.value.code == -5
    with pytest.raises(orb.OrbxError):
        orb.ORBextractor(1000, 1.2, 8, 20, 7)(synth(1, 200, 200))       # level 7 is 56x56 < 62
    k, d = ext(synth(2, 480, 640))                   # the handle stays usable
    assert len(k) > 100


def test_capacity_error(orb):
    import ctypes
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    img = synth(0, 375, 1242)
    kps = np.zeros(10, orb.KEYPOINT_DTYPE); desc = np.zeros((10, 32), np.uint8); n = ctypes.c_int()
    rc = ext._lib.orbx_extract(ext._h, img.ctypes.data, 1242, 375, 1242, kps.ctypes.data, desc.ctypes.data, 10, ctypes.byref(n))
    assert rc == -2 and n.value > 10 and b"cap" in ext._lib.orbx_last_error(ext._h)


def test_batch_equals_single_and_is_deterministic(orb):
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    frames = [synth(s, 375, 1242) for s in range(6)]
    singles = [ext(f) for f in frames]
    for rep in range(2):
        batch = ext.extract_batch(frames)
        for (k1, d1), (k2, d2) in zip(singles, batch):
            assert np.array_equal(k1, k2) and np.array_equal(d1, d2)


def test_device_resident_paths(orb):
    """Frames already in HBM: aligned stride (used in place) and odd stride (copied) give identical results."""
    import torch
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    frames = np.stack([synth(s, 375, 1242) for s in range(4)])
    ref = ext.extract_batch(list(frames))
    for pitch in (1242, 1280):
        buf = torch.zeros((4, 375, pitch), dtype=torch.uint8, device="cuda")
        buf[:, :, :1242] = torch.from_numpy(frames).cuda()
        torch.cuda.synchronize()
        ext.submit_device(buf.data_ptr(), 4, 1242, 375, pitch, 375 * pitch)
        kps, desc, n = ext.collect_view()
        for f in range(4):
            assert np.array_equal(kps[f, :n[f]], ref[f][0]) and np.array_equal(desc[f, :n[f]], ref[f][1]), (pitch, f)


def test_matcher_vs_oracle(orb, oracle_mod):
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    (_, dA), (_, dB) = ext.extract_batch([synth(0, 375, 1242), synth(1, 375, 1242)])
    m = orb.ORBmatcher(0.9, extractor=ext)
    rng = np.random.default_rng(1)
    rnd = rng.integers(0, 256, (777, 32), dtype=np.uint8)
    near = dA.copy(); flip = rng.integers(0, 256, len(near)); near[np.arange(len(near)), flip % 32] ^= 1 << (flip % 8).astype(np.uint8)
    for A, B in ((dA, dB), (dA, dA), (dA, near), (rnd, rnd[::-1].copy()), (dA[:1], dB), (dA, dB[:1]), (dA[:130], dB[:129])):
        for th in (orb.ORBmatcher.TH_LOW, orb.ORBmatcher.TH_HIGH):
            idx, d1, d2, acc = m.match(A, B, th, 0.9)
            oi, o1, o2, oa = oracle_mod.Oracle.match(A, B, th, 0.9)
            assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2) and np.array_equal(acc, oa)
    idx, d1, d2, acc = m.match(dA, dA, 50, 0.9)
    assert np.array_equal(idx[d1 == 0][:5], np.arange(len(dA))[d1 == 0][:5])      # self-match: distance 0 at the own (lowest) index
    idx, d1, d2, acc = m.match(dA, np.zeros((0, 32), np.uint8), 50, 0.9)
    assert (idx == -1).all() and (d1 == 256).all() and not acc.any()


def test_rotation_filter_vs_oracle(orb, oracle_mod):
    """mbCheckOrientation after the match (src/ORBmatcher.cc:610-620, 641-660, 2233-2274): bit-exact accept flags,
    histogram and the three selected bins, on real extractor output and on constructed tie / wrap-around cases."""
    from test_oracle_vs_ref import _rotation_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    (kA, dA), (kB, dB) = ext.extract_batch([synth(0, 375, 1242), synth(1, 375, 1242)])
    m = orb.ORBmatcher(0.9, True, extractor=ext)
    for A, angA, B, angB in ((dA, kA["angle"], dB, kB["angle"]), (dA, kA["angle"], dA, kA["angle"])):
        for th in (orb.ORBmatcher.TH_LOW, orb.ORBmatcher.TH_HIGH):
            idx, d1, d2, acc, hist, top3 = m.match_oriented(A, angA, B, angB, th, 0.9)
            oi, o1, o2, oa = oracle_mod.Oracle.match(A, B, th, 0.9)
            ka, oh, ot = oracle_mod.Oracle.rotation_filter(oi, oa, angA, angB)
            assert np.array_equal(idx, oi) and np.array_equal(acc, ka) and np.array_equal(hist, oh) and np.array_equal(top3, ot)
            assert acc.sum() == hist[top3[top3 >= 0]].sum()
    for idx, acc0, a, b in _rotation_cases():
        acc, hist, top3 = ext.rotation_filter(idx, acc0, a, b)
        ka, oh, ot = oracle_mod.Oracle.rotation_filter(idx, acc0, a, b)
        assert np.array_equal(acc, ka) and np.array_equal(hist, oh) and np.array_equal(top3, ot)
    with pytest.raises(orb.OrbxError):            # an accepted match must point into B
        ext.rotation_filter(np.array([5], np.int32), np.array([1], np.uint8), np.zeros(1, np.float32), np.zeros(3, np.float32))


@pytest.mark.parametrize("seed,shape,params", [(3, (375, 1242), (2000, 1.2, 8, 20, 7)), (5, (240, 400), (600, 1.2, 4, 20, 7))])
def test_stereo_matches_vs_oracle(orb, oracle_mod, seed, shape, params):
    """Frame::ComputeStereoMatches (src/Frame.cc:849-1038) on the device-resident results of two extractors:
    mvuRight, mvDepth bit-identical, vDescIndex identical to the oracle port (itself pinned to the reference's code)."""
    from test_oracle_vs_ref import _stereo_inputs
    L, R, kL, dL, kR, dR, t, pyrL, pyrR = _stereo_inputs(oracle_mod, seed, shape[0], shape[1], params)
    eL, eR = orb.ORBextractor(*params), orb.ORBextractor(*params)
    gkL, gdL = eL(L); gkR, gdR = eR(R)
    assert np.array_equal(gdL, dL) and np.array_equal(gdR, dR) and kps_equal_exact(gkL, kL) and kps_equal_exact(gkR, kR)
    exp = oracle_mod.Oracle.stereo_matches(kL, dL, kR, dR, t["scale"], t["inv_scale"], pyrL, pyrR, 386.1448)
    ur, dp, di, kept = eL.stereo_match(eR, 386.1448)
    assert kept == exp["kept"] and len(ur) == len(kL)
    assert np.array_equal(ur.view(np.uint32), exp["u_right"].view(np.uint32))
    assert np.array_equal(dp.view(np.uint32), exp["depth"].view(np.uint32))
    assert np.array_equal(di, exp["desc_index"])
    # frames of a batch: left = frame 1 of a two-frame batch, right = frame 0 of another
    eL.extract_batch([R, L]); eR.extract_batch([R, L])
    ur2, dp2, di2, kept2 = eL.stereo_match(eR, 386.1448, frame_left=1, frame_right=0)
    assert kept2 == kept and np.array_equal(ur2, ur) and np.array_equal(dp2, dp) and np.array_equal(di2, di)
    with pytest.raises(orb.OrbxError):
        eL.stereo_match(orb.ORBextractor(params[0], 1.2, params[2] - 1, 20, 7), 386.1448)


@pytest.mark.parametrize("channels,rgb", [(3, True), (3, False), (4, True), (4, False)])
def test_color_input_vs_oracle(orb, oracle_mod, channels, rgb):
    """Colour frames as Tracking::GrabImageRGBD receives them: device-side cvtColor (src/Tracking.cc:459-472) + extraction
    equal the oracle's grey conversion (checked against cv2 in tests/test_oracle_cvprim.py) + extraction."""
    import ctypes
    rng = np.random.default_rng(41)
    base = np.stack([synth(60 + c, 300, 501) for c in range(channels)], axis=2)        # odd width: ragged last quad
    img = np.ascontiguousarray(np.clip(base.astype(np.int16) + rng.integers(-3, 4, base.shape), 0, 255).astype(np.uint8))
    L = oracle_mod.Oracle.lib()
    gray = np.zeros(img.shape[:2], np.uint8)
    L.cvp_cvt_gray_u8(ctypes.c_void_p(img.ctypes.data), img.shape[1], img.shape[0], img.strides[0], channels, int(rgb),
                      ctypes.c_void_p(gray.ctypes.data), gray.strides[0])
    params = (1000, 1.2, 6, 20, 7)
    ext, o = orb.ORBextractor(*params), oracle_mod.Oracle(*params)
    kps, desc = ext.extract_color(img, rgb=rgb)
    ok, od = o(gray)
    assert np.array_equal(ext.pyramid_level(0), gray)
    assert kps_equal_exact(kps, ok) and np.array_equal(kps["angle"], ok["angle"]) and np.array_equal(desc, od)


def test_search_by_projection_vs_oracle(orb, oracle_mod):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1958-2102): projections, windows and
    Hamming distances on the GPU + the sequential claim walk on the host give the oracle's mvpMapPoints assignment and nmatches."""
    from test_oracle_vs_ref import _projection_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    m = orb.ORBmatcher(0.9, True, extractor=ext)
    for case in _projection_cases(oracle_mod):
        exp, nexp = oracle_mod.search_by_projection_port(case)
        got, n = m.SearchByProjection(case)
        assert n == nexp and np.array_equal(got, exp)
        no_ori = dict(case, check_orientation=False)
        exp2, nexp2 = oracle_mod.search_by_projection_port(no_ori)
        got2, n2 = m.SearchByProjection(no_ori)
        assert n2 == nexp2 and np.array_equal(got2, exp2) and n2 >= n
    empty = dict(case, world_pos=np.zeros((0, 3), np.float32), mp_desc=np.zeros((0, 32), np.uint8), valid=np.zeros(0, np.uint8),
                 nobs=np.zeros(0, np.int32), last_octave=np.zeros(0, np.int32), last_angle=np.zeros(0, np.float32))
    got, n = m.SearchByProjection(empty)
    assert n == 0 and (got == -1).all()


def test_search_for_initialization_vs_oracle(orb, oracle_mod):
    """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:780-895): window candidates + Hamming distances on the GPU and the
    vMatchedDistance / vnMatches21 bookkeeping replayed on the host give the oracle's vnMatches12, vbPrevMatched and nmatches."""
    from test_oracle_vs_ref import _initialization_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    m = orb.ORBmatcher(0.9, True, extractor=ext)
    for case in _initialization_cases(oracle_mod):
        for c in (case, dict(case, check_orientation=False)):
            em, ep, en = oracle_mod.search_for_initialization_port(c)
            gm, gp, gn = m.SearchForInitialization(c)
            assert gn == en and np.array_equal(gm, em) and np.array_equal(gp.view(np.uint32), ep.view(np.uint32))
    empty = dict(case, xy2=np.zeros((0, 2), np.float32), oct2=np.zeros(0, np.int32), ang2=np.zeros(0, np.float32), desc2=np.zeros((0, 32), np.uint8))
    gm, gp, gn = m.SearchForInitialization(empty)
    assert gn == 0 and (gm == -1).all() and np.array_equal(gp, case["prev_xy"])


def test_search_local_points_vs_oracle(orb, oracle_mod):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:418-502): windows, gates and Hamming distances on the GPU and
    the claim bookkeeping replayed on the host give the oracle's map point per feature and nmatches."""
    from test_oracle_vs_ref import _local_points_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    m = orb.ORBmatcher(0.8, True, extractor=ext)
    for case in _local_points_cases(oracle_mod):
        exp, nexp = oracle_mod.search_local_points_port(case)
        got, n = m.SearchLocalPoints(case)
        assert n == nexp and np.array_equal(got, exp)
    empty = dict(case, proj=np.zeros((0, 3), np.float32), view_cos=np.zeros(0, np.float32), level=np.zeros(0, np.int32),
                 mp_desc=np.zeros((0, 32), np.uint8), valid=np.zeros(0, np.uint8), nobs=np.zeros(0, np.int32))
    got, n = m.SearchLocalPoints(empty)
    assert n == 0 and (got == -1).all()


def test_search_by_bow_vs_oracle(orb, oracle_mod, tmp_path):
    """ORBmatcher::SearchByBoW(pKF, F, ..) (src/ORBmatcher.cc:532-663): pair distances inside the shared vocabulary nodes on the GPU and
    the skip-if-matched walk on the host give the oracle's matches; node ids once from the toy quantiser and once from the product's own
    vocabulary transform (Frame::ComputeBoW, levelsup 4 as the reference calls it) on a synthetic tree."""
    from test_oracle_vs_ref import _bow_match_cases, _voc_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    m = orb.ORBmatcher(0.7, True, extractor=ext)
    _, d = ext(synth(2, 375, 1242))
    path, k, L = _voc_cases(tmp_path, d)[0]
    voc = orb.ORBVocabulary(ext)
    assert voc.loadFromTextFile(path)
    node_fns = (None, lambda desc: voc.transform_each(np.ascontiguousarray(desc), max(L - 2, 0))[1])
    for fn in node_fns:
        for case in _bow_match_cases(oracle_mod, fn):
            for c in (case, dict(case, check_orientation=False)):
                exp, nexp = oracle_mod.search_by_bow_port(c)
                got, n = m.SearchByBoW(c)
                assert n == nexp and np.array_equal(got, exp)
                exp, nexp = oracle_mod.search_by_bow_kf_port(c)                  # SearchByBoW(pKF1, pKF2, ..), :897-1030
                got, n = m.SearchByBoWKeyFrames(c)
                assert n == nexp and np.array_equal(got, exp)
    empty = dict(case, f_angle=np.zeros(0, np.float32), f_desc=np.zeros((0, 32), np.uint8), f_nodes=np.zeros(0, np.int32),
                 f_off=np.zeros(1, np.int32), f_feats=np.zeros(0, np.int32))
    got, n = m.SearchByBoW(empty)
    assert n == 0 and len(got) == 0


def test_vocabulary_transform_vs_oracle(orb, oracle_mod, tmp_path):
    """Frame::ComputeBoW (src/Frame.cc:778-785): ORBVocabulary::loadFromTextFile + transform on the GPU against the oracle
    port (pinned to the reference's DBoW2): words, nodes, weights per descriptor; BowVector values bit-identical."""
    from test_oracle_vs_ref import _voc_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    _, d = ext(synth(2, 375, 1242))
    for path, k, L in _voc_cases(tmp_path, d):
        voc = orb.ORBVocabulary(ext)
        assert voc.loadFromTextFile(path)
        o = oracle_mod.OracleVocabulary(path)
        assert voc.info() == dict(k=o.k, L=o.levels, nodes=o.nodes, words=o.words)
        for lu in (4, 2, 0):
            for x, y in zip(voc.transform_each(d, lu), o.transform_each(d, lu)):
                assert np.array_equal(x, y)
            tg, to = voc.transform(d, lu), o.transform(d, lu)
            assert np.array_equal(tg[1].view(np.uint64), to[1].view(np.uint64))
            for x, y in zip(tg, to):
                assert np.array_equal(x, y)
        e = voc.transform(np.zeros((0, 32), np.uint8), 4)
        assert all(len(a) == 0 for a in e[:3])
    bad = tmp_path / "bad.txt"
    bad.write_text("99 3 0 0\n0 1 " + " ".join(["0"] * 32) + " 1.0")
    with pytest.raises(orb.OrbxError):
        orb.ORBVocabulary(ext).loadFromTextFile(str(bad))


def test_distinctive_descriptors_vs_oracle(orb, oracle_mod):
    """MapPoint::ComputeDistinctiveDescriptors batched over map points (src/MapPoint.cc:242-306): representative row and
    its median identical to the oracle, on constructed cases and on tracks built from real descriptors."""
    from test_oracle_vs_ref import _distinctive_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    for desc, off in (_distinctive_cases(23), _distinctive_cases(99)):
        best, med = ext.distinctive_descriptors(desc, off)
        ob, om_ = oracle_mod.Oracle.distinctive_descriptors(desc, off)
        assert np.array_equal(best, ob) and np.array_equal(med, om_)
    # "map points" = groups of descriptors of one frame that are close to each other, 3000 of them
    _, d = ext(synth(0, 375, 1242))
    rng = np.random.default_rng(5)
    sizes = rng.integers(1, 40, 3000)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    rows = d[rng.integers(0, len(d), off[-1])]
    best, med = ext.distinctive_descriptors(rows, off)
    ob, om_ = oracle_mod.Oracle.distinctive_descriptors(rows, off)
    assert np.array_equal(best, ob) and np.array_equal(med, om_)
    with pytest.raises(orb.OrbxError):
        ext.distinctive_descriptors(np.zeros((1100, 32), np.uint8), [0, 1100])


def test_full_size_properties(orb, oracle_mod):
    """BASELINE configs 3 and 5 at full size: one frame each against the oracle (seconds on the CPU),
    and size-independent properties on the batch: determinism, self-match, level-major ordering, bounds."""
    for (h, w), params in (((1080, 1920), (5000, 1.2, 8, 20, 7)), ((2160, 3840), (10000, 1.2, 12, 20, 7))):
        ext = orb.ORBextractor(*params)
        o = oracle_mod.Oracle(*params)
        img = synth(0, h, w)
        bad, total = check_frame(ext, o, img, "%dx%d" % (w, h))
        print("%dx%d descriptor mismatch %d/%d" % (w, h, bad, total))
        kps, desc = ext(img)
        k2, d2 = ext(img)
        assert np.array_equal(kps, k2) and np.array_equal(desc, d2)                       # idempotent
        assert (np.diff(kps["octave"]) >= 0).all()                                           # level-major order
        sc = ext.GetScaleFactors()
        lx, ly = kps["x"] / sc[kps["octave"]], kps["y"] / sc[kps["octave"]]
        for l in range(params[2]):
            lw, lh = ext.level_size(l)
            sel = kps["octave"] == l
            assert sel.sum() <= ext.mnFeaturesPerLevel[l] + 3
            assert (lx[sel] >= 19 - 1e-3).all() and (lx[sel] <= lw - 19).all() and (ly[sel] >= 19 - 1e-3).all() and (ly[sel] <= lh - 19).all()
        m = orb.ORBmatcher(0.9, extractor=ext)
        idx, d1, _, _ = m.match(desc, desc, 100, 0.9)
        assert (d1 == 0).all() and (idx <= np.arange(len(desc))).all()


def test_cpp_adapter_drop_in(orb, oracle_mod, tmp_path):
    """The C++ adapter with the reference's class surface (multimot_track_b200/adapter), driven like
    Frame::ExtractORB (src/Frame.cc:621), gives the oracle's keypoints, descriptors and mvImagePyramid."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "multimot_track_b200", "adapter", "adapter_demo")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(exe)])
    img = synth(3, 375, 1242)
    raw, out = tmp_path / "gray.raw", tmp_path / "out.bin"
    raw.write_bytes(img.tobytes())
    params = (2000, 1.2, 8, 20, 7)
    subprocess.check_call([exe, str(raw), "1242", "375"] + [str(p) for p in params] + [str(out)])
    blob = out.read_bytes()
    n = int(np.frombuffer(blob, np.int32, 1)[0])
    kps = np.frombuffer(blob, orb.KEYPOINT_DTYPE, n, 4)
    desc = np.frombuffer(blob, np.uint8, n * 32, 4 + 28 * n).reshape(n, 32)
    o = oracle_mod.Oracle(*params)
    okps, odesc = o(img)
    assert kps_equal_exact(kps, okps) and np.array_equal(kps["angle"], okps["angle"])
    assert np.array_equal(desc, odesc)
    off = 4 + 60 * n
    for l in range(8):
        w, h = (int(v) for v in np.frombuffer(blob, np.int32, 2, off)); off += 8
        padded = np.frombuffer(blob, np.uint8, (w + 38) * (h + 38), off).reshape(h + 38, w + 38); off += (w + 38) * (h + 38)
        assert np.array_equal(padded, o.level_padded(l)), l


def test_tma_and_plain_staging_agree(orb, oracle_mod):
    """The TMA-staged kernels (cp.async.bulk.tensor) and their plain staging fallbacks (ORBX_OPT_TMA_STAGING = 0), the persistent TMA
    variant of the FAST kernel (ORBX_OPT_FAST_TMA) and the copied-input path (ORBX_OPT_COPY_INPUT) all reproduce the oracle bit for bit."""
    import torch
    img = synth(11, 375, 1242)
    o = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    check_frame(ext, o, img, "tma")
    ext.set_option(ext.OPT_FAST_TMA, 1)
    check_frame(ext, o, img, "fast-tma")
    ext.set_option(ext.OPT_FAST_TMA, 0)
    ext.set_option(ext.OPT_TMA_STAGING, 0)
    check_frame(ext, o, img, "plain")
    ext.set_option(ext.OPT_TMA_STAGING, 1)
    check_frame(ext, o, img, "tma again")
    with pytest.raises(orb.OrbxError):
        ext.set_option(99, 1)
    # device-resident, aligned frames: in place by default, copied with ORBX_OPT_COPY_INPUT (then the caller's buffer may be
    # overwritten after the collect without changing what level 0 reads back)
    okps, odesc = o(img)
    buf = torch.zeros((1, 375, 1280), dtype=torch.uint8, device="cuda")
    buf[0, :, :1242] = torch.from_numpy(img).cuda()
    torch.cuda.synchronize()
    for copy in (0, 1):
        ext.set_option(ext.OPT_COPY_INPUT, copy)
        ext.submit_device(buf.data_ptr(), 1, 1242, 375, 1280, 375 * 1280)
        kps, desc, n = ext.collect_view()
        assert kps_equal_exact(kps[0, :n[0]], okps) and np.array_equal(desc[0, :n[0]], odesc)
    buf.zero_()
    torch.cuda.synchronize()
    assert np.array_equal(ext.pyramid_level(0), img)
