"""oracle/cv2_restatement.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python restatement of ORBextractor::operator() (src/ORBextractor.cc:1046-1109 of the reference) on the real OpenCV primitives
of the cv2 4.13 wheel -- the only OpenCV in this image, and the OpenCV semantics the oracle is pinned to:

  ComputePyramid           :1111-1136   cv2.resize INTER_LINEAR level from level, cv2.copyMakeBorder REFLECT_101
  per-cell FAST            :765-829     cv2.FastFeatureDetector (TYPE_9_16, nonmaxSuppression) on every ~30-pixel cell, iniThFAST then minThFAST
  DistributeOctTree        :539-763     the oracle port's orbo_distribute (canonical tie-break; pinned to the reference's compiled function)
  IC_Angle                 :77-104      numpy, all keypoints of a level at once; cv2.fastAtan2's polynomial restated in float32
  GaussianBlur             :1089-1090   cv2.GaussianBlur(7x7, 2, 2, BORDER_REFLECT_101) on a copy of the level
  computeOrbDescriptor     :108-147     numpy gathers, float32 rotation, round-half-even

Two uses: an independent cross-check of the C oracle (tests/test_oracle_vs_ref.py, where cv2 is
importable), and the second CPU baseline of bench.py: T single-threaded processes (cv2.setNumThreads(1)), one frame at a time each.
"""
import ctypes
import os

import numpy as np

HALF_PATCH, EDGE, PATCH = 15, 19, 31
_PATTERN = None


def _pattern():
    global _PATTERN
    if _PATTERN is None:
        txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "orb_pattern.inc")).read()
        import re
        vals = [int(t) for t in re.findall(r"-?\d+", re.sub(r"/\*.*?\*/", "", txt, flags=re.S))]
        _PATTERN = np.array(vals[:1024], np.int32).reshape(512, 2)        # bit_pattern_31_ (:150-408): 256 pairs of (x, y)
    return _PATTERN


def fast_atan2_deg(y, x):
    """cv::fastAtan2 on float32 arrays, no fused multiply-adds: every product / sum rounds to float32."""
    f = np.float32
    y = y.astype(f); x = x.astype(f)
    scale = f(180.0 / np.pi)
    p1, p3, p5, p7 = (f(0.9997878412794807) * scale, f(-0.3258083974640975) * scale, f(0.1555786518463281) * scale, f(-0.04432655554792128) * scale)
    ax, ay = np.abs(x), np.abs(y)
    eps = f(2.2204460492503131e-16)
    swap = ax < ay
    num = np.where(swap, ax, ay); den = np.where(swap, ay, ax) + eps
    c = (num / den).astype(f)
    c2 = (c * c).astype(f)
    a = ((((p7 * c2).astype(f) + p5).astype(f) * c2).astype(f) + p3).astype(f)
    a = (((a * c2).astype(f) + p1).astype(f) * c).astype(f)
    a = np.where(swap, (f(90.0) - a).astype(f), a)
    a = np.where(x < 0, (f(180.0) - a).astype(f), a)
    a = np.where(y < 0, (f(360.0) - a).astype(f), a)
    return a.astype(f)


class Cv2Extractor:
    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        import cv2
        from .oracle import Oracle
        self.cv2 = cv2
        self.params = (int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        self.nlevels = int(nlevels)
        t = Oracle(*self.params).tables()              # constructor tables (:410-470): float / double arithmetic restated in C, pinned to the reference
        self.scale, self.inv_scale, self.nfeat, self.umax = t["scale"], t["inv_scale"], t["nfeat"], t["umax"]
        self.det_ini = cv2.FastFeatureDetector_create(int(iniThFAST), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self.det_min = cv2.FastFeatureDetector_create(int(minThFAST), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        self.lib = Oracle.lib()
        # circular patch as (v, u) offsets of IC_Angle (:77-101): rows |v| <= 15, |u| <= umax[|v|]
        vs, us = [], []
        for v in range(-HALF_PATCH, HALF_PATCH + 1):
            d = int(self.umax[abs(v)])
            for u in range(-d, d + 1):
                vs.append(v); us.append(u)
        self.pv, self.pu = np.array(vs, np.int32), np.array(us, np.int32)
        pat = _pattern()
        self.px, self.py = pat[:, 0].astype(np.float32), pat[:, 1].astype(np.float32)
        self.pyramid, self.blurred, self.candidates = [], [], []

    def _level_sizes(self, h, w):
        return [(int(np.rint(np.float32(w) * self.inv_scale[l])), int(np.rint(np.float32(h) * self.inv_scale[l]))) for l in range(self.nlevels)]

    def compute_pyramid(self, image):
        cv2 = self.cv2
        self.pyramid, self.padded = [], []
        for l, (w, h) in enumerate(self._level_sizes(*image.shape)):
            lvl = image if l == 0 else cv2.resize(self.pyramid[l - 1], (w, h), interpolation=cv2.INTER_LINEAR)
            pad = cv2.copyMakeBorder(lvl, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)       # mvImagePyramid's parent buffer (:1126-1132)
            self.padded.append(pad)
            self.pyramid.append(pad[EDGE:-EDGE, EDGE:-EDGE])
        return self.pyramid

    def _cells_fast(self, img):
        """ComputeKeyPointsOctTree's cell loop (:771-829): candidates (x, y, score) relative to (16, 16), reference order."""
        rows, cols = img.shape
        minB, maxBX, maxBY = EDGE - 3, cols - EDGE + 3, rows - EDGE + 3
        width, height = np.float32(maxBX - minB), np.float32(maxBY - minB)
        nCols, nRows = int(width / np.float32(30)), int(height / np.float32(30))
        wCell, hCell = int(np.ceil(width / np.float32(nCols))), int(np.ceil(height / np.float32(nRows)))
        out = []
        for i in range(nRows):
            iniY = minB + i * hCell
            maxY = iniY + hCell + 6
            if iniY >= maxBY - 3:
                continue
            maxY = min(maxY, maxBY)
            for j in range(nCols):
                iniX = minB + j * wCell
                maxX = iniX + wCell + 6
                if iniX >= maxBX - 6:
                    continue
                maxX = min(maxX, maxBX)
                cell = img[iniY:maxY, iniX:maxX]
                kps = self.det_ini.detect(cell)
                if not kps:
                    kps = self.det_min.detect(cell)
                for k in kps:
                    out.append((int(k.pt[0]) + j * wCell, int(k.pt[1]) + i * hCell, int(k.response)))
        return np.array(out, np.int32).reshape(-1, 3), (minB, maxBX, minB, maxBY)

    def _distribute(self, cand, bounds, N):
        minX, maxX, minY, maxY = bounds
        c = np.ascontiguousarray(cand, np.int32)
        out = np.zeros((len(c) + 8, 3), np.int32)
        n = self.lib.orbo_distribute(c.ctypes.data, len(c), minX, maxX, minY, maxY, int(N), out.ctypes.data, len(out))
        return out[:n]

    def _angles(self, img, xs, ys):
        """IC_Angle for all keypoints of a level: integer moments over the circular patch, fastAtan2."""
        patch = img[(ys[:, None] + self.pv[None, :]), (xs[:, None] + self.pu[None, :])].astype(np.int32)
        m10 = (patch * self.pu[None, :]).sum(1)
        m01 = (patch * self.pv[None, :]).sum(1)
        return fast_atan2_deg(m01.astype(np.float32), m10.astype(np.float32))

    def _descriptors(self, blur, xs, ys, angles):
        f = np.float32
        ang = (angles.astype(f) * f(np.pi / 180.0)).astype(f)            # (float)(CV_PI / 180.f), :112
        a, b = np.cos(ang.astype(np.float64)).astype(f)[:, None], np.sin(ang.astype(np.float64)).astype(f)[:, None]
        px, py = self.px[None, :], self.py[None, :]
        r = np.rint(((px * b).astype(f) + (py * a).astype(f)).astype(f)).astype(np.int32)       # cvRound: half to even
        c = np.rint(((px * a).astype(f) - (py * b).astype(f)).astype(f)).astype(np.int32)
        v = blur[ys[:, None] + r, xs[:, None] + c]
        bits = (v[:, 0::2] < v[:, 1::2]).astype(np.uint8)               # bit k of byte j = I(p[16j+2k]) < I(p[16j+2k+1])
        return np.packbits(bits.reshape(len(xs), 32, 8), axis=2, bitorder="little").reshape(len(xs), 32)

    def __call__(self, image):
        from .oracle import KEYPOINT_DTYPE
        cv2 = self.cv2
        image = np.ascontiguousarray(image, np.uint8)
        self.compute_pyramid(image)
        kps_all, desc_all = [], []
        self.candidates, self.blurred = [], []
        for l, img in enumerate(self.pyramid):
            cand, bounds = self._cells_fast(img)
            self.candidates.append(cand)
            sel = self._distribute(cand, bounds, self.nfeat[l]) if len(cand) else cand
            if len(sel) == 0:
                self.blurred.append(None)
                continue
            xs, ys = sel[:, 0] + bounds[0], sel[:, 1] + bounds[2]
            ang = self._angles(img, xs, ys)
            blur = cv2.GaussianBlur(img.copy(), (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
            self.blurred.append(blur)
            desc_all.append(self._descriptors(blur, xs, ys, ang))
            k = np.zeros(len(sel), KEYPOINT_DTYPE)
            sc = self.scale[l]
            k["x"] = xs.astype(np.float32) * sc if l else xs
            k["y"] = ys.astype(np.float32) * sc if l else ys
            k["size"] = np.float32(int(np.float32(PATCH) * sc)); k["angle"] = ang; k["response"] = sel[:, 2]
            k["octave"] = l; k["class_id"] = -1
            kps_all.append(k)
        if not kps_all:
            return np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        return np.concatenate(kps_all), np.concatenate(desc_all)


def _worker(args):
    """One process of the cv2 CPU baseline: its own extractor, single-threaded OpenCV, its share of the frames.  Returns seconds."""
    import time
    params, frames, reps = args
    import cv2
    cv2.setNumThreads(1)
    ext = Cv2Extractor(*params)
    ext(frames[0])
    t0 = time.perf_counter()
    n = 0
    for _ in range(reps):
        for f in frames:
            k, _ = ext(f)
            n += len(k)
    return time.perf_counter() - t0, n


def _noop(_):
    return 0


def extract_many_processes(params, frames, processes, reps=1):
    """frames [F,H,W] split round-robin over `processes` single-threaded worker processes; returns (wall seconds, keypoints).
    Forks: call it from a process without CUDA / helper threads (bench.py runs this module as a subprocess)."""
    import multiprocessing as mp
    import time
    parts = [frames[p::processes] for p in range(processes)]
    parts = [p for p in parts if len(p)]
    ctx = mp.get_context("fork")
    with ctx.Pool(len(parts)) as pool:
        pool.map(_noop, range(len(parts)))                       # workers are up before the clock starts
        t0 = time.perf_counter()
        res = pool.map(_worker, [(params, p, reps) for p in parts])
        wall = time.perf_counter() - t0
    return wall, sum(r[1] for r in res)


def main():
    """python -m oracle.cv2_restatement H W nfeatures nlevels nframes processes min_seconds -> one JSON line (the cv2 CPU leg)."""
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import cv2
    from multimot_track_b200.synth import frame_pool
    H, W, nfeat, nlev, nframes, procs = (int(v) for v in sys.argv[1:7])
    min_s = float(sys.argv[7])
    params = (nfeat, 1.2, nlev, 20, 7)
    frames = frame_pool(H, W, nframes, 0)
    reps = 1
    wall, nkp = extract_many_processes(params, frames, procs, reps)
    while wall < min_s and reps < 256:
        reps *= 2
        wall, nkp = extract_many_processes(params, frames, procs, reps)
    print(json.dumps({"value": nframes * reps / wall, "unit": "frames/s", "cores": procs, "frames": nframes * reps, "wall_s": wall,
                      "keypoints": nkp, "cv2": cv2.__version__, "ms_per_frame_per_core": 1e3 * wall * procs / (nframes * reps)}))


if __name__ == "__main__":
    main()
