"""Parity on the configurations bench.py actually times ("what's weak" 1-2): every frame of the 64-frame
synthetic pool through the batched entry points at the benchmark's batch sizes, with the benchmark's six handles in
flight, compared with the CPU oracle frame by frame -- all seven cv::KeyPoint fields and the descriptors
(src/ORBextractor.cc:1046-1109).  The launch shapes that depend on the batch size (octree split launch, resize chunking,
orientation / descriptor CTAs per frame) are therefore compared with the oracle, not only with themselves.  Plus the
matcher at BASELINE.json configs[3]'s size on extracted descriptors (src/ORBmatcher.cc:574-605) and config 1's
consecutive KITTI frames."""
import concurrent.futures as cf
import os

import numpy as np
import pytest

from conftest import angle_diff_rad, desc_bit_mismatch, kps_equal_exact

pytestmark = pytest.mark.gpu

ANGLE_TOL_RAD = 1e-4
DESC_MISMATCH_TOL = 1e-4
POOL = 64


@pytest.fixture(scope="module")
def orb():
    import torch
    assert torch.cuda.is_available(), "these tests need the B200"
    import multimot_track_b200
    multimot_track_b200.load_library()
    return multimot_track_b200


def oracle_pool(oracle_mod, params, frames):
    """Oracle results of every frame, one Oracle instance per worker thread (the C port releases the GIL)."""
    workers = min(len(frames), os.cpu_count() or 1)

    def work(chunk):
        o = oracle_mod.Oracle(*params)
        out = []
        for i in chunk:
            k, d = o(frames[i])
            out.append((i, k.copy(), d.copy()))
        return out
    res = [None] * len(frames)
    with cf.ThreadPoolExecutor(workers) as ex:
        for part in ex.map(work, [range(w, len(frames), workers) for w in range(workers)]):
            for i, k, d in part:
                res[i] = (k, d)
    return res


def compare(kps, desc, n, idx, ref, tag):
    bad_total = bits_total = 0
    for f, i in enumerate(idx):
        k, d = kps[f, :n[f]], desc[f, :n[f]]
        rk, rd = ref[i]
        assert len(k) == len(rk), "%s frame %d: %d keypoints, oracle %d" % (tag, i, len(k), len(rk))
        assert kps_equal_exact(k, rk), "%s frame %d: keypoint fields" % (tag, i)
        assert angle_diff_rad(k["angle"], rk["angle"]).max(initial=0) <= ANGLE_TOL_RAD, "%s frame %d: angles" % (tag, i)
        bad, total = desc_bit_mismatch(d, rd)
        bad_total += bad; bits_total += total
    assert bad_total <= DESC_MISMATCH_TOL * bits_total, "%s: %d of %d descriptor bits differ" % (tag, bad_total, bits_total)
    return bad_total, bits_total


def run_pool(orb, oracle_mod, H, W, params, batch, nsteps, nhandles=6, host_steps=1):
    """bench.py's pipelined loop (step s on handle s % n, collected right before the handle is reused), checked."""
    import torch
    from multimot_track_b200.synth import frame_pool
    pool = frame_pool(H, W, POOL, 0)
    ref = oracle_pool(oracle_mod, params, pool)
    pitch = (W + 63) // 64 * 64
    dpool = torch.zeros((POOL, H, pitch), dtype=torch.uint8, device="cuda")
    dpool[:, :, :W] = torch.from_numpy(pool).cuda()
    # the batches wrap around the pool: keep a second copy behind it so that a batch is one contiguous device range
    dpool = torch.cat([dpool, dpool], 0).contiguous()
    torch.cuda.synchronize()
    handles = [orb.ORBextractor(*params, max_width=W, max_height=H, max_batch=batch) for _ in range(nhandles)]
    stride = max(1, POOL // max(1, nsteps)) if batch >= POOL else batch
    pending = [None] * nhandles
    bad = bits = frames_checked = 0

    def collect(i):
        nonlocal bad, bits, frames_checked
        kps, desc, n = handles[i].collect_view()
        b, t = compare(kps, desc, n, pending[i], ref, "%dx%d batch %d" % (W, H, batch))
        bad += b; bits += t; frames_checked += len(pending[i])
        pending[i] = None
    for s in range(nsteps):
        i = s % nhandles
        if pending[i] is not None:
            collect(i)
        first = (s * stride) % POOL
        handles[i].submit_device(dpool[first].data_ptr(), batch, W, H, pitch, H * pitch)
        pending[i] = [(first + j) % POOL for j in range(batch)]
    for i in range(nhandles):
        if pending[i] is not None:
            collect(i)
    # the host-buffer entry point (bench.py's e2e leg) at the same batch size
    for s in range(host_steps):
        idx = [(7 + s * batch + j) % POOL for j in range(batch)]
        handles[0].submit_host([pool[i] for i in idx])
        kps, desc, n = handles[0].collect_view()
        b, t = compare(kps, desc, n, idx, ref, "%dx%d host batch %d" % (W, H, batch))
        bad += b; bits += t; frames_checked += batch
    print("%dx%d batch %d: %d frames checked against the oracle, %d / %d descriptor bits differ" % (W, H, batch, frames_checked, bad, bits))
    return ref, pool


def test_k1_pool_batch32_six_handles(orb, oracle_mod):
    """BASELINE.json configs[1] as bench.py runs it: 1242x375, 2000 features, batch 32, six handles; 12 steps = every pool frame 6 times."""
    run_pool(orb, oracle_mod, 375, 1242, (2000, 1.2, 8, 20, 7), 32, 12)


def test_k2_pool_batch64(orb, oracle_mod):
    """configs[2]: 1920x1080, 5000 features, batch 64 (the whole pool per batch), two batches with different first frames, six handles."""
    run_pool(orb, oracle_mod, 1080, 1920, (5000, 1.2, 8, 20, 7), 64, 2)


def test_k4_pool_batch8(orb, oracle_mod):
    """configs[4]: 3840x2160, 12 levels, 10000 features, batch 8, six handles; 8 steps = all 64 pool frames."""
    run_pool(orb, oracle_mod, 2160, 3840, (10000, 1.2, 12, 20, 7), 8, 8, host_steps=1)


def _fit_rows(d, rows):
    """Extractor output padded (cyclically) / cropped to exactly `rows` descriptors."""
    reps = -(-rows // len(d))
    return np.ascontiguousarray(np.concatenate([d] * reps)[:rows])


def test_matcher_5000x5000_extracted(orb, oracle_mod):
    """configs[3] at full size on extractor output of pool frames s, s+1 (1920x1080, 5000 features -> 5007 rows, cropped to 5000),
    TH_LOW and TH_HIGH, ratio 0.9: idx / d1 / d2 / accept equal the oracle's scan; accepted counts reported."""
    from multimot_track_b200.synth import value_noise_frame
    ext = orb.ORBextractor(5000, 1.2, 8, 20, 7)
    descs = [_fit_rows(ext(value_noise_frame(s, 1080, 1920))[1], 5000) for s in (0, 1, 2)]
    m = orb.ORBmatcher(0.9, extractor=ext)
    for a, b in ((0, 1), (1, 2), (0, 0)):
        for th in (orb.ORBmatcher.TH_LOW, orb.ORBmatcher.TH_HIGH):
            idx, d1, d2, acc = m.match(descs[a], descs[b], th, 0.9)
            oi, o1, o2, oa = oracle_mod.Oracle.match(descs[a], descs[b], th, 0.9, threads=os.cpu_count() or 1)
            assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2) and np.array_equal(acc, oa), (a, b, th)
            print("matcher 5000x5000 frames %d/%d TH %d: %d accepted" % (a, b, th, int(acc.sum())))
    # chunk seams of k_match_partial: sizes around the tile / chunk boundaries
    for nA, nB in ((5000, 4999), (4097, 5000), (1, 5000), (5000, 129)):
        idx, d1, d2, acc = m.match(descs[0][:nA], descs[1][:nB], 100, 0.9)
        oi, o1, o2, oa = oracle_mod.Oracle.match(descs[0][:nA], descs[1][:nB], 100, 0.9, threads=os.cpu_count() or 1)
        assert np.array_equal(idx, oi) and np.array_equal(d1, o1) and np.array_equal(d2, o2) and np.array_equal(acc, oa), (nA, nB)


def test_kitti_consecutive_frame_matching(orb, oracle_mod, golden_kitti, kitti_frames):
    """Config 1: kitti_sample frames 0..4 with kitti03.yaml's parameters, frame i matched against frame i+1; descriptors from the GPU
    extractor, result against the golden made from the reference-compiled descriptors (scripts/make_golden.py) and the oracle live."""
    ext = orb.ORBextractor(4000, 1.2, 8, 20, 7)
    out = ext.extract_batch(kitti_frames)
    for f in range(5):
        assert kps_equal_exact(out[f][0], golden_kitti["kps_f%d_n4000" % f])
        assert np.array_equal(out[f][1], golden_kitti["desc_f%d_n4000" % f])
    m = orb.ORBmatcher(0.9, extractor=ext)
    for f in range(4):
        for th in (50, 100):
            idx, d1, d2, acc = m.match(out[f][1], out[f + 1][1], th, 0.9)
            g = golden_kitti["match_f%d_th%d" % (f, th)].astype(np.int32)
            assert np.array_equal(idx, g[:, 0]) and np.array_equal(d1, g[:, 1]) and np.array_equal(d2, g[:, 2]) and np.array_equal(acc.astype(np.int32), g[:, 3])
            oi, o1, o2, oa = oracle_mod.Oracle.match(out[f][1], out[f + 1][1], th, 0.9)
            assert np.array_equal(idx, oi) and np.array_equal(acc, oa)
            print("kitti %d->%d TH %d: %d of %d accepted" % (f, f + 1, th, int(acc.sum()), len(idx)))
