"""Synthetic frames of SURVEY.md section 8d: the value-noise generator whose FAST
corner density is close to the KITTI sample's.  Needs cv2 (in the image) only for the
INTER_CUBIC up-sampling of the coarse grid; CRC32 anchors in tests/golden pin it."""
import numpy as np


def value_noise_frame(seed, height, width):
    import cv2
    rng = np.random.default_rng(seed)
    rows, cols = -(-height // 12) + 2, -(-width // 12) + 2
    coarse = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
    up = cv2.resize(coarse, (12 * cols, 12 * rows), interpolation=cv2.INTER_CUBIC)[12:12 + height, 12:12 + width]
    noise = rng.integers(-4, 5, (height, width), dtype=np.int16)
    return np.clip(up.astype(np.int16) + noise, 0, 255).astype(np.uint8)


def frame_pool(height, width, count=64, first_seed=0):
    return np.stack([value_noise_frame(first_seed + i, height, width) for i in range(count)])


def uniform_noise_frame(seed, height, width):
    return np.random.default_rng(seed).integers(0, 256, (height, width), dtype=np.uint8)


def stereo_pair(seed, h, w, max_disp=60.0):
    """A rectified stereo pair for Frame::ComputeStereoMatches tests: the right image is the left one resampled
    with a smooth positive disparity field d(x, y) in [2, max_disp] (right(x) = left(x + d)), plus +-2 of fresh noise.
    The images are 16 px wider internally so the shifted samples exist."""
    import cv2
    left_wide = value_noise_frame(seed, h, w + int(max_disp) + 8)
    rng = np.random.default_rng(seed + 7919)
    coarse = rng.random((h // 64 + 3, w // 64 + 3)).astype(np.float32)
    disp = cv2.resize(coarse, (w, h), interpolation=cv2.INTER_CUBIC)
    disp = 2.0 + (max_disp - 2.0) * np.clip(disp, 0, 1)
    xs = np.arange(w, dtype=np.float32)[None, :].repeat(h, 0)
    ys = np.arange(h, dtype=np.float32)[:, None].repeat(w, 1)
    right = cv2.remap(left_wide, xs + disp, ys, cv2.INTER_LINEAR)
    right = np.clip(right.astype(np.int16) + rng.integers(-2, 3, (h, w), dtype=np.int16), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left_wide[:, :w]), right


def write_synthetic_vocabulary(path, k=10, levels=4, seed=0, scoring=0, weighting=0, seeds=None, stop_fraction=0.02):
    """A complete k-ary vocabulary tree in the ORBvoc text format (`k L scoring weighting`, then per node
    `parent isLeaf d0 .. d31 weight`, TemplatedVocabulary.h:1338-1418 / saveToTextFile :1423-1448), written in the
    creation order of DBoW2's HKmeans (children of a node consecutive, depth first).  Node descriptors are their
    parent's with a shrinking number of random bit flips (children of the root: `seeds`, real descriptors, if given);
    leaf weights are idf-like positive doubles, a few of them 0 (stopped words)."""
    rng = np.random.default_rng(seed)
    rows = []

    def grow(parent_id, parent_desc, level):
        ids = []
        for c in range(k):
            if level == 1 and seeds is not None:
                d = np.array(seeds[rng.integers(0, len(seeds))], np.uint8).copy()
            else:
                d = parent_desc.copy()
                for f in rng.integers(0, 256, max(2, 96 >> level)):
                    d[f % 32] ^= np.uint8(1 << (f % 8))
            leaf = level == levels
            w = 0.0 if (leaf and rng.random() < stop_fraction) else (float(rng.random() * 9 + 0.5) if leaf else 0.0)
            rows.append((parent_id, int(leaf), d, w))
            ids.append(len(rows))
        if level < levels:
            for nid in ids:
                grow(nid, rows[nid - 1][2], level + 1)

    # DBoW2 numbers the k children of a node consecutively and then recurses into each of them
    grow(0, rng.integers(0, 256, 32, dtype=np.uint8), 1)
    # no newline after the last node: DBoW2's loader loops on !eof() and would read an empty line as one more node
    with open(path, "w") as f:
        f.write("%d %d %d %d\n" % (k, levels, scoring, weighting))
        f.write("\n".join("%d %d %s %s" % (parent_id, leaf, " ".join(str(int(v)) for v in d), repr(w)) for parent_id, leaf, d, w in rows))
    return len(rows) + 1


def projection_case(seed, kps, desc, scale, th=15.0, mono=False, forward=0.0, distractors=None):
    """A TrackWithMotionModel-shaped input for ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono):
    the last frame = the given keypoints / descriptors with random depths, back-projected to world points through
    KITTI-like intrinsics (Tcw_last = identity); the current frame = those points seen from a slightly moved camera,
    +-1.5 px of noise, a few descriptor bits flipped, octaves occasionally off by one, shuffled, plus distractor features;
    mvuRight consistent with the depth for two thirds of them.  10 % of the last-frame points carry no usable map point,
    a third of the map points have no observation yet (temporal points)."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    fx, fy, cx, cy, bf = 718.856, 718.856, 607.1928, 185.2157, 386.1448
    w, h = 1242.0, 375.0
    z = rng.uniform(4.0, 60.0, n).astype(np.float32)
    xw = ((kps["x"] - np.float32(cx)) * z / np.float32(fx)).astype(np.float32)
    yw = ((kps["y"] - np.float32(cy)) * z / np.float32(fy)).astype(np.float32)
    world = np.stack([xw, yw, z], 1).astype(np.float32)
    ang = np.deg2rad(rng.uniform(-1.5, 1.5, 3))
    Rx = np.array([[1, 0, 0], [0, np.cos(ang[0]), -np.sin(ang[0])], [0, np.sin(ang[0]), np.cos(ang[0])]])
    Ry = np.array([[np.cos(ang[1]), 0, np.sin(ang[1])], [0, 1, 0], [-np.sin(ang[1]), 0, np.cos(ang[1])]])
    Rz = np.array([[np.cos(ang[2]), -np.sin(ang[2]), 0], [np.sin(ang[2]), np.cos(ang[2]), 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rx @ Ry @ Rz
    T[:3, 3] = [rng.uniform(-0.2, 0.2), rng.uniform(-0.05, 0.05), -forward + rng.uniform(-0.1, 0.1)]
    Tc = T.astype(np.float32)
    pc = (Tc[:3, :3].astype(np.float64) @ world.T.astype(np.float64) + Tc[:3, 3:4].astype(np.float64)).T
    u = fx * pc[:, 0] / pc[:, 2] + cx + rng.uniform(-1.5, 1.5, n)
    v = fy * pc[:, 1] / pc[:, 2] + cy + rng.uniform(-1.5, 1.5, n)
    cd = desc.copy()
    for i in range(n):
        for f in rng.integers(0, 256, rng.integers(0, 25)):
            cd[i, f % 32] ^= np.uint8(1 << (f % 8))
    coct = np.clip(kps["octave"] + rng.choice([-1, 0, 0, 0, 0, 1], n), 0, len(scale) - 1).astype(np.int32)
    cang = ((kps["angle"] + rng.normal(0, 4, n)) % 360).astype(np.float32)
    ur = np.where(rng.random(n) < 0.66, u - bf / pc[:, 2] + rng.uniform(-1, 1, n), -1.0)
    if distractors is not None and len(distractors):
        m = len(distractors)
        u = np.concatenate([u, rng.uniform(20, w - 20, m)]); v = np.concatenate([v, rng.uniform(20, h - 20, m)])
        cd = np.concatenate([cd, distractors]); coct = np.concatenate([coct, rng.integers(0, len(scale), m).astype(np.int32)])
        cang = np.concatenate([cang, rng.uniform(0, 360, m).astype(np.float32)]); ur = np.concatenate([ur, np.full(m, -1.0)])
    perm = rng.permutation(len(u))
    return dict(cam=np.array([fx, fy, cx, cy, bf, bf / fx, 0.0, w, 0.0, h], np.float32), Tcw_cur=Tc.reshape(-1).copy(),
                Tcw_last=np.eye(4, dtype=np.float32).reshape(-1).copy(), world_pos=world, mp_desc=np.ascontiguousarray(desc),
                valid=(rng.random(n) < 0.9).astype(np.uint8), nobs=np.where(rng.random(n) < 0.33, 0, rng.integers(1, 9, n)).astype(np.int32),
                last_octave=kps["octave"].astype(np.int32), last_angle=kps["angle"].astype(np.float32),
                cur_xy=np.stack([u, v], 1).astype(np.float32)[perm], cur_octave=coct[perm].astype(np.int32), cur_angle=cang[perm].astype(np.float32),
                cur_uright=ur.astype(np.float32)[perm], cur_desc=np.ascontiguousarray(cd[perm]), scale=np.asarray(scale, np.float32),
                th=float(th), mono=bool(mono), check_orientation=True)


def initialization_case(seed, kps, desc, window=100, nnratio=0.9, shift=(6.0, -3.0), distractors=None):
    """Monocular initialisation input for ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize):
    F2 = F1's keypoints moved by `shift` + noise with a few descriptor bits flipped, near-duplicates so that matches get stolen,
    shuffled, plus distractors; vbPrevMatched = F1's positions (as Tracking::MonocularInitialization seeds it, src/Tracking.cc:2594-2596)."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    x2 = kps["x"] + np.float32(shift[0]) + rng.uniform(-2, 2, n).astype(np.float32)
    y2 = kps["y"] + np.float32(shift[1]) + rng.uniform(-2, 2, n).astype(np.float32)
    d2 = desc.copy()
    for i in range(n):
        for f in rng.integers(0, 256, rng.integers(0, 30)):
            d2[i, f % 32] ^= np.uint8(1 << (f % 8))
    o2 = kps["octave"].astype(np.int32).copy()
    a2 = ((kps["angle"] + rng.normal(0, 5, n)) % 360).astype(np.float32)
    dup = rng.choice(n, n // 6, replace=False)                     # a second feature near the true one with a similar descriptor
    x2 = np.concatenate([x2, x2[dup] + rng.uniform(-8, 8, len(dup)).astype(np.float32)])
    y2 = np.concatenate([y2, y2[dup] + rng.uniform(-8, 8, len(dup)).astype(np.float32)])
    dd = d2[dup].copy()
    for i in range(len(dup)):
        for f in rng.integers(0, 256, rng.integers(0, 12)):
            dd[i, f % 32] ^= np.uint8(1 << (f % 8))
    d2 = np.concatenate([d2, dd]); o2 = np.concatenate([o2, o2[dup]]); a2 = np.concatenate([a2, a2[dup]])
    if distractors is not None and len(distractors):
        m = len(distractors)
        x2 = np.concatenate([x2, rng.uniform(20, 1220, m).astype(np.float32)]); y2 = np.concatenate([y2, rng.uniform(20, 355, m).astype(np.float32)])
        d2 = np.concatenate([d2, distractors]); o2 = np.concatenate([o2, np.zeros(m, np.int32)]); a2 = np.concatenate([a2, rng.uniform(0, 360, m).astype(np.float32)])
    perm = rng.permutation(len(x2))
    return dict(cam=np.array([718.856, 718.856, 607.1928, 185.2157, 386.1448, 0.5372, 0.0, 1242.0, 0.0, 375.0], np.float32),
                xy1=np.stack([kps["x"], kps["y"]], 1).astype(np.float32), oct1=kps["octave"].astype(np.int32), ang1=kps["angle"].astype(np.float32),
                desc1=np.ascontiguousarray(desc), xy2=np.stack([x2, y2], 1).astype(np.float32)[perm], oct2=o2[perm].astype(np.int32),
                ang2=a2[perm].astype(np.float32), desc2=np.ascontiguousarray(d2[perm]),
                prev_xy=np.stack([kps["x"], kps["y"]], 1).astype(np.float32), window=int(window), nnratio=float(nnratio), check_orientation=True)


def local_points_case(seed, kps, desc, scale, th=3.0, nnratio=0.8, distractors=None):
    """Input for ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:418-502, called by Tracking::SearchLocalPoints):
    the frame's features (kps / desc of an extractor call, some with a right coordinate, some already holding a map point) and local
    map points that project near them -- 1.4 points per feature on average so that claims collide, predicted levels at the feature's
    octave or one above, viewing cosines on both sides of 0.998, descriptors with a few flipped bits."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    w, h, bf = 1242.0, 375.0, 386.1448
    x = kps["x"].astype(np.float32); y = kps["y"].astype(np.float32)
    octave = kps["octave"].astype(np.int32)
    d = np.ascontiguousarray(desc)
    if distractors is not None and len(distractors):
        m = len(distractors)
        x = np.concatenate([x, rng.uniform(20, w - 20, m).astype(np.float32)]); y = np.concatenate([y, rng.uniform(20, h - 20, m).astype(np.float32)])
        octave = np.concatenate([octave, rng.integers(0, len(scale), m).astype(np.int32)]); d = np.concatenate([d, distractors])
    nf = len(x)
    depth = rng.uniform(4.0, 40.0, nf)
    uright = np.where(rng.random(nf) < 0.6, x - bf / depth, -1.0).astype(np.float32)
    feat_obs = np.where(rng.random(nf) < 0.15, rng.integers(0, 6, nf), -1).astype(np.int32)        # -1: no map point, 0: unobserved one
    src = np.concatenate([rng.permutation(n), rng.choice(n, int(0.4 * n))])                         # the feature each map point comes from
    npnt = len(src)
    sc = np.asarray(scale, np.float32)
    level = np.clip(octave[src] + rng.choice([0, 0, 1, 1, 2, -1], npnt), 0, len(sc) - 1).astype(np.int32)
    spread = sc[level] * rng.choice([0.5, 2.0, 5.0], npnt)
    proj = np.stack([x[src] + rng.uniform(-1, 1, npnt) * spread, y[src] + rng.uniform(-1, 1, npnt) * spread,
                     np.where(uright[src] > 0, uright[src] + rng.uniform(-1, 1, npnt) * spread * 1.5, x[src] - bf / 10.0)], 1).astype(np.float32)
    mpd = d[src].copy()
    for i in range(npnt):
        for f in rng.integers(0, 256, rng.integers(0, 40)):
            mpd[i, f % 32] ^= np.uint8(1 << (f % 8))
    view_cos = rng.choice(np.array([0.9, 0.99, 0.998, 0.9981, 0.9995, 1.0], np.float32), npnt).astype(np.float32)
    perm = rng.permutation(nf)
    inv = np.empty(nf, np.int64); inv[perm] = np.arange(nf)
    return dict(cam=np.array([718.856, 718.856, 607.1928, 185.2157, bf, 0.5372, 0.0, w, 0.0, h], np.float32), proj=proj, view_cos=view_cos,
                level=level, mp_desc=np.ascontiguousarray(mpd), valid=(rng.random(npnt) < 0.9).astype(np.uint8),
                nobs=np.where(rng.random(npnt) < 0.3, 0, rng.integers(1, 9, npnt)).astype(np.int32),
                xy=np.stack([x, y], 1).astype(np.float32)[perm], octave=octave[perm].astype(np.int32), uright=uright[perm],
                desc=np.ascontiguousarray(d[perm]), feat_obs=feat_obs[perm], scale=sc, th=float(th), nnratio=float(nnratio))


def feature_vector(nodes):
    """DBoW2::FeatureVector of per-feature node ids, flattened in map order: (node ids ascending, offsets, feature indices ascending
    within a node), the layout orbx_voc_bow returns."""
    nodes = np.asarray(nodes)
    order = np.argsort(nodes, kind="stable")
    ids, counts = np.unique(nodes, return_counts=True)
    return ids.astype(np.int32), np.concatenate([[0], np.cumsum(counts)]).astype(np.int32), order.astype(np.int32)


def _toy_nodes(desc):
    """A stand-in for the vocabulary node of a descriptor that is stable under a few bit flips: quantised popcounts of byte groups."""
    pc = np.unpackbits(desc, axis=1).reshape(len(desc), 4, 64).sum(2)
    return ((pc[:, 0] // 6) * 36 + (pc[:, 1] // 6) * 6 + pc[:, 2] // 6).astype(np.int32)


def bow_match_case(seed, kps, desc, nnratio=0.7, distractors=None, node_fn=_toy_nodes):
    """Input for ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches, ..) (src/ORBmatcher.cc:532-663): the key frame = one extractor
    call (85 % of its features hold a good map point); the frame = its descriptors with a few flipped bits, near-duplicates (so that
    claims collide inside a node), shuffled, plus distractors; node ids per feature from `node_fn` (a vocabulary transform or the toy
    quantiser above), flattened like DBoW2::FeatureVector."""
    rng = np.random.default_rng(seed)
    n = len(kps)
    fd = desc.copy()
    for i in range(n):
        for f in rng.integers(0, 256, rng.integers(0, 24)):
            fd[i, f % 32] ^= np.uint8(1 << (f % 8))
    fa = ((kps["angle"] + rng.normal(0, 5, n)) % 360).astype(np.float32)
    dup = rng.choice(n, n // 5, replace=False)
    dd = fd[dup].copy()
    for i in range(len(dup)):
        for f in rng.integers(0, 256, rng.integers(0, 6)):
            dd[i, f % 32] ^= np.uint8(1 << (f % 8))
    fd = np.concatenate([fd, dd]); fa = np.concatenate([fa, ((fa[dup] + rng.normal(0, 30, len(dup))) % 360).astype(np.float32)])
    if distractors is not None and len(distractors):
        fd = np.concatenate([fd, distractors]); fa = np.concatenate([fa, rng.uniform(0, 360, len(distractors)).astype(np.float32)])
    perm = rng.permutation(len(fd))
    fd = np.ascontiguousarray(fd[perm]); fa = fa[perm].astype(np.float32)
    kd = np.ascontiguousarray(desc)
    kn, ko, kf = feature_vector(node_fn(kd))
    fn_, fo, ff = feature_vector(node_fn(fd))
    return dict(kf_angle=kps["angle"].astype(np.float32), kf_desc=kd, kf_valid=(rng.random(n) < 0.85).astype(np.uint8), kf_nodes=kn, kf_off=ko, kf_feats=kf,
                f_angle=fa, f_desc=fd, f_nodes=fn_, f_off=fo, f_feats=ff, nnratio=float(nnratio), check_orientation=True,
                f_valid=(rng.random(len(fd)) < 0.8).astype(np.uint8))      # used when the second side is a key frame (SearchByBoW(pKF1, pKF2))
