"""GPU tests of the frame-sharded dispatcher (orbx_pool_*) and of the robustness fixes of round 2:
garbage levels of invalid map points, search windows with more than 512 candidates, out-of-range / NaN angles in the rotation
histogram, and the current-device side effect of the entry points."""
import ctypes

import numpy as np
import pytest

from conftest import desc_bit_mismatch, kps_equal_exact

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def orb():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    import multimot_track_b200
    multimot_track_b200.load_library()
    return multimot_track_b200


def synth(seed, h, w):
    from multimot_track_b200.synth import value_noise_frame
    return value_noise_frame(seed, h, w)


def _check_shards(shards, frames_idx, ref, nshards):
    seen = 0
    for g, (first, kps, desc, n) in enumerate(shards):
        assert first == seen, "shards are contiguous blocks in frame order"
        for f in range(len(n)):
            rk, rd = ref[frames_idx[first + f]]
            assert kps_equal_exact(kps[f, :n[f]], rk) and np.array_equal(kps[f, :n[f]]["angle"], rk["angle"])
            assert np.array_equal(desc[f, :n[f]], rd)
        seen += len(n)
    assert seen == len(frames_idx) and len(shards) == nshards


@pytest.mark.parametrize("devices", [None, [0], [0, 0], [0, 0, 0]])
def test_pool_shards_equal_oracle(orb, oracle_mod, devices):
    """Host frames through orbx_pool_submit_host: contiguous blocks per worker (two / three workers share the one GPU of the test box),
    several tickets in flight, collected out of order; every frame equals the oracle.  Then device-resident shards."""
    import torch
    params = (1000, 1.2, 8, 20, 7)
    H, W = 480, 640
    frames = [synth(s, H, W) for s in range(13)]
    o = oracle_mod.Oracle(*params)
    ref = [tuple(a.copy() for a in o(f)) for f in frames]
    pool = orb.ExtractorPool(*params, devices=devices, depth=3, max_width=W, max_height=H, max_batch=8)
    G = pool.nshards
    assert G == (len(devices) if devices else 1) and pool.depth == 3
    assert [pool.shard_range(13, g) for g in range(G)] == [(13 * g // G, 13 * (g + 1) // G - 13 * g // G) for g in range(G)]
    sets = [list(range(13)), [12, 3, 5, 7, 1], [4], list(range(6, 13)) + list(range(6))]
    t0 = pool.submit_host([frames[i] for i in sets[0]])
    t1 = pool.submit_host([frames[i] for i in sets[1]])
    t2 = pool.submit_host([frames[i] for i in sets[2]])             # one frame: the other workers get empty shards
    with pytest.raises(orb.OrbxError) as e:                         # depth tickets are uncollected
        pool.submit_host([frames[0]])
    assert e.value.code == -6
    _check_shards(pool.collect(t1), sets[1], ref, G)
    _check_shards(pool.collect(t0), sets[0], ref, G)
    t3 = pool.submit_host([frames[i] for i in sets[3]])
    _check_shards(pool.collect(t2), sets[2], ref, G)
    _check_shards(pool.collect(t3), sets[3], ref, G)
    with pytest.raises(orb.OrbxError):
        pool.collect(t3)                                            # a ticket is collected once
    # device-resident shards: per-worker pointers into one device array (aligned pitch: read in place)
    pitch = 640
    dev = torch.from_numpy(np.stack(frames)).cuda()
    torch.cuda.synchronize()
    for rep in range(4):                                            # more submits than depth: the ring wraps
        ptrs, counts, idx = [], [], []
        for g in range(G):
            first, cnt = pool.shard_range(12, g)
            ptrs.append(dev[first + rep % 2].data_ptr()); counts.append(cnt)
            idx += list(range(first + rep % 2, first + rep % 2 + cnt))
        t = pool.submit_device(ptrs, counts, W, H, pitch, H * pitch)
        _check_shards(pool.collect(t), idx, ref, G)
    assert pool.launch_count() > 0
    pool.close()


def test_pool_create_errors(orb):
    with pytest.raises(orb.OrbxError):
        orb.ExtractorPool(1000, 1.2, 8, 20, 7, devices=[99])
    with pytest.raises(orb.OrbxError):
        orb.ExtractorPool(1000, 1.2, 99, 20, 7)


def test_entry_points_restore_the_current_device(orb):
    """Every entry point runs on its handle's device and leaves the caller's current device alone (one GPU here: the call must at
    least not fail or change it)."""
    import torch
    before = torch.cuda.current_device()
    ext = orb.ORBextractor(500, 1.2, 4, 20, 7, device_id=0)
    ext(synth(1, 240, 320))
    assert torch.cuda.current_device() == before


def test_local_points_ignore_the_level_of_invalid_points(orb, oracle_mod):
    """mnTrackScaleLevel is uninitialised for map points outside the frustum (mbTrackInView false): whatever is in level[] for them
    must neither change the result nor fault."""
    from test_oracle_vs_ref import _local_points_cases
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    m = orb.ORBmatcher(0.8, True, extractor=ext)
    case = _local_points_cases(oracle_mod)[0]
    exp, nexp = oracle_mod.search_local_points_port(case)
    bad = dict(case)
    lvl = case["level"].copy()
    inv = np.flatnonzero(case["valid"] == 0)
    assert len(inv) > 10
    lvl[inv[0::3]] = -1; lvl[inv[1::3]] = 0x7fffffff; lvl[inv[2::3]] = -0x80000000
    bad["level"] = lvl
    got, n = m.SearchLocalPoints(bad)
    assert n == nexp and np.array_equal(got, exp)
    ext(synth(0, 240, 320))                                         # no sticky fault on the context


def test_windows_with_more_than_512_candidates(orb, oracle_mod):
    """The reference has no cap on GetFeaturesInArea's result.  Dense frames with wide windows (> 512 candidates in one window, the
    shared-memory staging size) take the second run with global staging and still equal the oracle: all three windowed matchers."""
    from multimot_track_b200.synth import initialization_case, local_points_case, projection_case
    o = oracle_mod.Oracle(8000, 1.2, 8, 20, 7)
    k, d = o(synth(0, 375, 1242))
    _, d2 = oracle_mod.Oracle(2000, 1.2, 8, 20, 7)(synth(1, 375, 1242))
    sc = o.tables()["scale"]
    ext = orb.ORBextractor(8000, 1.2, 8, 20, 7)
    m = orb.ORBmatcher(0.9, True, extractor=ext)
    ic = initialization_case(1, k, d, 400, 0.9, (6.0, -3.0), d2[:600])
    em, ep, en = oracle_mod.search_for_initialization_port(ic)
    gm, gp, gn = m.SearchForInitialization(ic)
    assert gn == en and np.array_equal(gm, em) and np.array_equal(gp.view(np.uint32), ep.view(np.uint32))
    pc = projection_case(5, k, d, sc, 150.0, False, 0.0, d2[:700])
    exp, nexp = oracle_mod.search_by_projection_port(pc)
    got, n = m.SearchByProjection(pc)
    assert n == nexp and np.array_equal(got, exp)
    lc = local_points_case(3, k, d, sc, 40.0, 0.8, d2[:700])
    exp, nexp = oracle_mod.search_local_points_port(lc)
    got, n = m.SearchLocalPoints(lc)
    assert n == nexp and np.array_equal(got, exp)
    # the cases really overflow the shared-memory staging: count the densest window on the host
    xy2 = ic["xy2"]; c0 = ic["prev_xy"][ic["oct1"] == 0]
    dens = max(int(((np.abs(xy2[:, 0] - u) < 400) & (np.abs(xy2[:, 1] - v) < 400) & (ic["oct2"] == 0)).sum()) for u, v in c0[:50])
    assert dens > 512, dens


def test_rotation_histogram_with_invalid_angles(orb, oracle_mod):
    """Angles outside [0, 360) or NaN (uninitialised keypoints) must not index the 30-bin histogram out of bounds: such matches
    are pruned, the others behave as before."""
    ext = orb.ORBextractor(2000, 1.2, 8, 20, 7)
    n = 400
    rng = np.random.default_rng(3)
    idx = rng.integers(0, n, n).astype(np.int32)
    acc = np.ones(n, np.uint8)
    a = rng.uniform(0, 360, n).astype(np.float32); b = (a[idx] if False else rng.uniform(0, 360, n)).astype(np.float32)
    good_acc, good_hist, good_top = ext.rotation_filter(idx, acc, a, b)
    ka, oh, ot = oracle_mod.Oracle.rotation_filter(idx, acc, a, b)
    assert np.array_equal(good_acc, ka) and np.array_equal(good_hist, oh)
    a2 = a.copy()
    a2[::7] = np.nan; a2[1::7] = 1e9; a2[2::7] = -1e9
    acc2, hist2, top2 = ext.rotation_filter(idx, acc, a2, b)
    bad = np.zeros(n, bool); bad[::7] = bad[1::7] = bad[2::7] = True
    assert not acc2[bad].any()                                       # pruned
    ok = ~bad
    ka2, oh2, ot2 = oracle_mod.Oracle.rotation_filter(idx[ok], acc[ok], a[ok], b)     # the valid matches alone
    assert np.array_equal(hist2, oh2) and np.array_equal(acc2[ok], ka2)


def test_compact_keypoints_are_lossless(orb, oracle_mod):
    """ORBX_OPT_COMPACT_KEYPOINTS: 12-byte records on the way to the host; orbx_expand_keypoints and orbx_collect rebuild the
    28-byte cv::KeyPoint exactly (all seven fields bit-identical to the full-record path and to the oracle), descriptors unchanged;
    the device-resident consumers (stereo association) keep working; the pool delivers the same."""
    params = (2000, 1.2, 8, 20, 7)
    frames = [synth(s, 375, 1242) for s in range(5)]
    o = oracle_mod.Oracle(*params)
    ref = [tuple(a.copy() for a in o(f)) for f in frames]
    ext = orb.ORBextractor(*params)
    full = ext.extract_batch(frames)
    ext.set_option(ext.OPT_COMPACT_KEYPOINTS, 1)
    comp = ext.extract_batch(frames)                                 # orbx_collect expands on the host
    for f in range(5):
        assert np.array_equal(full[f][0], comp[f][0]) and np.array_equal(full[f][1], comp[f][1])
        assert kps_equal_exact(comp[f][0], ref[f][0]) and np.array_equal(comp[f][0]["angle"], ref[f][0]["angle"])
    ext.submit_host(frames)
    with pytest.raises(orb.OrbxError):
        ext.collect_view()                                           # full-record views are not available in compact mode
    ck, cd, n = ext.collect_view_compact()
    assert ck.dtype.itemsize == 12
    for f in range(5):
        k = ext.expand_keypoints(ck[f, :n[f]])
        assert np.array_equal(k, full[f][0]) and np.array_equal(cd[f, :n[f]], full[f][1])
        assert (ck[f, :n[f]]["octave"] == full[f][0]["octave"]).all() and (ck[f, :n[f]]["response"] == full[f][0]["response"]).all()
    # other scale factor / level count: the expansion uses the handle's own tables
    p2 = (900, 1.37, 5, 15, 5)
    e2 = orb.ORBextractor(*p2)
    img = synth(9, 480, 640)
    a = e2(img)
    e2.set_option(e2.OPT_COMPACT_KEYPOINTS, 1)
    b = e2(img)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # stereo association still reads the full records that stay on the device
    L, R = __import__("multimot_track_b200.synth", fromlist=["stereo_pair"]).stereo_pair(3, 375, 1242)
    eL, eR = orb.ORBextractor(*params), orb.ORBextractor(*params)
    eL(L); eR(R)
    base = eL.stereo_match(eR, 386.1448)
    eL.set_option(eL.OPT_COMPACT_KEYPOINTS, 1); eR.set_option(eR.OPT_COMPACT_KEYPOINTS, 1)
    eL(L); eR(R)
    again = eL.stereo_match(eR, 386.1448)
    assert base[3] == again[3] and all(np.array_equal(x, y) for x, y in zip(base[:3], again[:3]))
    # pool: compact at creation, switched off and on again between tickets
    pool = orb.ExtractorPool(*params, devices=[0, 0], depth=2, max_width=1242, max_height=375, max_batch=4, compact_keypoints=True)
    for mode in (1, 0, 1):
        pool.set_option(orb.ORBextractor.OPT_COMPACT_KEYPOINTS, mode)
        shards = pool.collect(pool.submit_host(frames))
        for first, kps, desc, n in shards:
            assert kps.dtype.itemsize == (12 if mode else 28)
            for f in range(len(n)):
                k = pool.expand_keypoints(kps[f, :n[f]]) if mode else kps[f, :n[f]]
                assert np.array_equal(k, full[first + f][0]) and np.array_equal(desc[f, :n[f]], full[first + f][1])
    t = pool.submit_host(frames)
    with pytest.raises(orb.OrbxError):
        pool.set_option(orb.ORBextractor.OPT_COMPACT_KEYPOINTS, 0)   # a ticket is outstanding
    pool.collect(t)
    pool.close()


@pytest.mark.parametrize("params", [(6000, 1.2, 12, 20, 7), (3000, 2.0, 5, 20, 7)])
def test_4k_octree_cluster_and_fallback(orb, oracle_mod, params):
    """3840x2160: with 12 levels x1.2 every level's octree sort runs on a cluster of four CTAs (sort launch + tree launch); with a
    steep pyramid (x2.0, 5 levels) the small top level has no room for the cluster's counters in its scratch slot and the whole
    image takes the single-CTA path.  Both equal the oracle, as a single frame and inside a batch of three."""
    frames = [synth(s, 2160, 3840) for s in (21, 22, 23)]
    o = oracle_mod.Oracle(*params)
    ref = [tuple(a.copy() for a in o(f)) for f in frames]
    ext = orb.ORBextractor(*params)
    k, d = ext(frames[0])
    assert kps_equal_exact(k, ref[0][0]) and np.array_equal(k["angle"], ref[0][0]["angle"]) and np.array_equal(d, ref[0][1])
    for (k, d), (rk, rd) in zip(ext.extract_batch(frames), ref):
        assert kps_equal_exact(k, rk) and np.array_equal(d, rd)
