"""Python mirror of the reference's ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111).

Same constructor arguments, same call shape (`extractor(image, mask) -> keypoints,
descriptors`, the mask being ignored exactly like src/ORBextractor.cc:58), the six
getters Frame reads (src/Frame.cc:87-93) and `mvImagePyramid`.  All computation
happens in liborbx.so on the GPU; this file only marshals buffers.
"""
import ctypes

import numpy as np

from . import _lib

# binary layout of cv::KeyPoint / orbx_keypoint (28 bytes)
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])


# orbx_keypoint_compact (12 bytes): position in the keypoint's own pyramid level, octave, FAST response, angle
COMPACT_KEYPOINT_DTYPE = np.dtype([("x", "<u2"), ("y", "<u2"), ("octave", "u1"), ("response", "u1"), ("reserved", "<u2"), ("angle", "<f4")])


class _PyramidView:
    """`mvImagePyramid`: level images of the last processed frame, fetched from the GPU on access."""

    def __init__(self, owner):
        self._o = owner

    def __len__(self):
        return self._o.nlevels

    def __getitem__(self, level):
        return self._o.pyramid_level(level)


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1          # include/ORBextractor.h:49

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, device_id=-1,
                 max_width=0, max_height=0, max_batch=0):
        self._lib = _lib.load_library()
        cfg = _lib.OrbxConfig(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST),
                              int(max_width), int(max_height), int(max_batch), int(device_id))
        h = ctypes.c_void_p()
        rc = self._lib.orbx_create(ctypes.byref(cfg), ctypes.byref(h))
        if rc != 0:
            raise _lib.OrbxError(rc, (self._lib.orbx_last_error(None) or b"").decode())
        self._h = h
        self.nfeatures, self.nlevels = int(nfeatures), int(nlevels)
        self.scaleFactor = float(np.float32(scaleFactor))
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)
        L = self.nlevels
        self._tables = [np.zeros(L, np.float32) for _ in range(4)] + [np.zeros(L, np.int32)]
        self._ck(self._lib.orbx_get_tables(self._h, *[t.ctypes.data for t in self._tables]))
        self.mvImagePyramid = _PyramidView(self)
        self._last_batch = 0

    # -- lifetime -------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        return _lib.check(self._lib, self._h, rc)

    # -- getters (include/ORBextractor.h:63-85) -------------------------------------
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return self.scaleFactor

    def GetScaleFactors(self):
        return self._tables[0].copy()

    def GetInverseScaleFactors(self):
        return self._tables[1].copy()

    def GetScaleSigmaSquares(self):
        return self._tables[2].copy()

    def GetInverseScaleSigmaSquares(self):
        return self._tables[3].copy()

    @property
    def mnFeaturesPerLevel(self):
        return self._tables[4].copy()

    def max_keypoints(self, width, height):
        return self._ck(self._lib.orbx_max_keypoints(self._h, int(width), int(height)))

    # -- operator() ------------------------------------------------------------
    @staticmethod
    def _as_gray(image):
        a = np.asarray(image)
        if a.ndim != 2 or a.dtype != np.uint8:
            raise AssertionError("image.type() == CV_8UC1")          # src/ORBextractor.cc:1053
        if a.strides[1] != 1:
            a = np.ascontiguousarray(a)
        return a

    def __call__(self, image, mask=None):
        """ORBextractor::operator()(image, mask, keypoints, descriptors): returns
        (keypoints[KEYPOINT_DTYPE], descriptors[n,32] uint8).  Empty image -> empty outputs."""
        if image is None or np.asarray(image).size == 0:
            return np.zeros(0, KEYPOINT_DTYPE), np.zeros((0, 32), np.uint8)
        a = self._as_gray(image)
        h, w = a.shape
        cap = self.max_keypoints(w, h)
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = ctypes.c_int(0)
        self._ck(self._lib.orbx_extract(self._h, a.ctypes.data, w, h, a.strides[0], kps.ctypes.data, desc.ctypes.data,
                                        cap, ctypes.byref(n)))
        self._last_batch = 1
        return kps[:n.value], desc[:n.value]

    def extract_batch(self, frames):
        """Host frames (list of HxW uint8 arrays or one [F,H,W] array) -> list of (keypoints, descriptors)."""
        arrs = [self._as_gray(f) for f in frames]
        h, w = arrs[0].shape
        stride = arrs[0].strides[0]
        for a in arrs:
            if a.shape != (h, w) or a.strides[0] != stride:
                raise ValueError("all frames of a batch need the same shape and stride")
        F = len(arrs)
        cap = self.max_keypoints(w, h)
        kps = np.zeros((F, cap), KEYPOINT_DTYPE)
        desc = np.zeros((F, cap, 32), np.uint8)
        n = np.zeros(F, np.int32)
        ptrs = (ctypes.c_void_p * F)(*[a.ctypes.data for a in arrs])
        self._ck(self._lib.orbx_extract_batch(self._h, ptrs, F, w, h, stride, kps.ctypes.data, desc.ctypes.data, cap,
                                              n.ctypes.data))
        self._last_batch = F
        return [(kps[f, :n[f]], desc[f, :n[f]]) for f in range(F)]

    # -- device-resident batches (torch tensors or raw device pointers) --------------
    def submit_device(self, ptr, nframes, width, height, stride_bytes, frame_stride_bytes):
        self._ck(self._lib.orbx_submit_device(self._h, ctypes.c_void_p(int(ptr)), int(nframes), int(width), int(height),
                                              int(stride_bytes), int(frame_stride_bytes)))
        self._last_batch = int(nframes)

    def submit_host(self, arrs):
        arrs = [self._as_gray(f) for f in arrs]
        h, w = arrs[0].shape
        F = len(arrs)
        ptrs = (ctypes.c_void_p * F)(*[a.ctypes.data for a in arrs])
        self._keepalive = arrs
        self._ck(self._lib.orbx_submit_host(self._h, ptrs, F, w, h, arrs[0].strides[0]))
        self._last_batch = F

    def collect_view(self):
        """Wait for the pending batch; zero-copy numpy views of the pinned result buffers
        (valid until the next submit): kps [F,cap], desc [F,cap,32], n [F]."""
        pk, pd, pn, cap = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
        self._ck(self._lib.orbx_collect_view(self._h, ctypes.byref(pk), ctypes.byref(pd), ctypes.byref(pn), ctypes.byref(cap)))
        F, c = self._last_batch, cap.value
        kb = (ctypes.c_uint8 * (F * c * 28)).from_address(pk.value)
        db = (ctypes.c_uint8 * (F * c * 32)).from_address(pd.value)
        nb = (ctypes.c_int32 * F).from_address(pn.value)
        kps = np.frombuffer(kb, dtype=KEYPOINT_DTYPE).reshape(F, c)
        desc = np.frombuffer(db, dtype=np.uint8).reshape(F, c, 32)
        n = np.frombuffer(nb, dtype=np.int32)
        return kps, desc, n

    OPT_TMA_STAGING, OPT_FAST_TMA, OPT_COPY_INPUT, OPT_COMPACT_KEYPOINTS = 1, 2, 3, 4        # ORBX_OPT_* of include/orbx.h

    def collect_view_compact(self):
        """As collect_view with ORBX_OPT_COMPACT_KEYPOINTS set: kps [F,cap] of COMPACT_KEYPOINT_DTYPE."""
        pk, pd, pn, cap = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_int()
        self._ck(self._lib.orbx_collect_view_compact(self._h, ctypes.byref(pk), ctypes.byref(pd), ctypes.byref(pn), ctypes.byref(cap)))
        F, c = self._last_batch, cap.value
        kb = (ctypes.c_uint8 * (F * c * 12)).from_address(pk.value)
        db = (ctypes.c_uint8 * (F * c * 32)).from_address(pd.value)
        nb = (ctypes.c_int32 * F).from_address(pn.value)
        return (np.frombuffer(kb, dtype=COMPACT_KEYPOINT_DTYPE).reshape(F, c), np.frombuffer(db, dtype=np.uint8).reshape(F, c, 32),
                np.frombuffer(nb, dtype=np.int32))

    def expand_keypoints(self, ckps):
        """orbx_expand_keypoints: compact records -> cv::KeyPoint records (exact)."""
        c = np.ascontiguousarray(ckps, COMPACT_KEYPOINT_DTYPE)
        out = np.zeros(len(c), KEYPOINT_DTYPE)
        self._ck(self._lib.orbx_expand_keypoints(self._h, c.ctypes.data, len(c), out.ctypes.data))
        return out

    def set_option(self, option, value):
        self._ck(self._lib.orbx_set_option(self._h, int(option), int(value)))

    def set_profiling(self, on=True):
        self._ck(self._lib.orbx_set_profiling(self._h, int(bool(on))))

    def stage_ms(self):
        """Device milliseconds per stage of the last collected batch (profiling must be on): dict name -> ms."""
        ms = (ctypes.c_float * 16)()
        n = self._ck(self._lib.orbx_get_stage_ms(self._h, ms, 16))
        return {self._lib.orbx_stage_name(i).decode(): float(ms[i]) for i in range(n)}

    def sync(self):
        self._ck(self._lib.orbx_sync(self._h))

    @property
    def stream(self):
        return self._lib.orbx_stream(self._h)

    @property
    def launch_count(self):
        return int(self._lib.orbx_launch_count(self._h))

    # -- stage read-back ----------------------------------------------------------
    def level_size(self, level):
        w, h = ctypes.c_int(), ctypes.c_int()
        rc = self._lib.orbx_get_level_size(self._h, int(level), ctypes.byref(w), ctypes.byref(h))
        if rc != 0:
            raise _lib.OrbxError(rc, "no geometry yet / level out of range")
        return w.value, h.value

    def pyramid_level(self, level, frame=0, with_border=False):
        w, h = self.level_size(level)
        b = 19 if with_border else 0
        out = np.zeros((h + 2 * b, w + 2 * b), np.uint8)
        self._ck(self._lib.orbx_get_pyramid_level(self._h, int(frame), int(level), out.ctypes.data, out.strides[0], int(with_border)))
        return out

    def blurred_level(self, level, frame=0):
        w, h = self.level_size(level)
        out = np.zeros((h, w), np.uint8)
        self._ck(self._lib.orbx_get_blurred_level(self._h, int(frame), int(level), out.ctypes.data, out.strides[0]))
        return out

    def candidates(self, level, frame=0):
        """FAST candidates before the octree, reference order: int32 [n,3] = x, y (relative to (16,16)), score."""
        n = ctypes.c_int()
        self._ck(self._lib.orbx_get_candidates(self._h, int(frame), int(level), None, 0, ctypes.byref(n)))
        out = np.zeros((max(n.value, 1), 3), np.int32)
        self._ck(self._lib.orbx_get_candidates(self._h, int(frame), int(level), out.ctypes.data, n.value, ctypes.byref(n)))
        return out[:n.value]

    # -- matcher entry points share the handle (stream + scratch) -----------------------
    def match(self, descA, descB, th, ratio):
        A = np.ascontiguousarray(descA, np.uint8).reshape(-1, 32)
        B = np.ascontiguousarray(descB, np.uint8).reshape(-1, 32)
        nA, nB = len(A), len(B)
        idx = np.full(nA, -1, np.int32); d1 = np.full(nA, 256, np.int32); d2 = np.full(nA, 256, np.int32)
        acc = np.zeros(nA, np.uint8)
        rc = self._lib.orbx_match(self._h, A.ctypes.data, nA, B.ctypes.data, nB, int(th), float(ratio),
                                  idx.ctypes.data, d1.ctypes.data, d2.ctypes.data, acc.ctypes.data)
        self._ck(rc)
        return idx, d1, d2, acc.astype(bool)

    def rotation_filter(self, idx, accept, angleA, angleB):
        """mbCheckOrientation (src/ORBmatcher.cc:610-620, 641-660, 2233-2274) on the GPU.
        Returns (accept_after[nA] bool, hist[30], top3[3])."""
        idx = np.ascontiguousarray(idx, np.int32)
        acc = np.ascontiguousarray(accept, np.uint8).copy()
        a = np.ascontiguousarray(angleA, np.float32); b = np.ascontiguousarray(angleB, np.float32)
        if len(acc) != len(idx) or len(a) != len(idx):
            raise ValueError("idx, accept and angleA must have one entry per query")
        hist = np.zeros(30, np.int32); top3 = np.zeros(3, np.int32)
        rc = self._lib.orbx_rotation_filter(self._h, len(idx), idx.ctypes.data, acc.ctypes.data, a.ctypes.data, b.ctypes.data, len(b),
                                            hist.ctypes.data, top3.ctypes.data)
        self._ck(rc)
        return acc.astype(bool), hist, top3

    def stereo_match(self, right, bf, frame_left=0, frame_right=0):
        """Frame::ComputeStereoMatches (src/Frame.cc:849-1038): this extractor's last results are the left frame,
        `right` (another ORBextractor on the same device) holds the right one.
        Returns (mvuRight[n], mvDepth[n], vDescIndex[n], matches_kept)."""
        n0 = ctypes.c_int(0)                         # first call: how many left keypoints are there
        self._ck(self._lib.orbx_stereo_match(self._h, right._h, int(frame_left), int(frame_right), float(bf), None, None, None, 0, ctypes.byref(n0)))
        cap = max(n0.value, 1)
        ur = np.zeros(cap, np.float32); dp = np.zeros(cap, np.float32); di = np.zeros(cap, np.int32)
        n = ctypes.c_int(0)
        rc = self._lib.orbx_stereo_match(self._h, right._h, int(frame_left), int(frame_right), float(bf), ur.ctypes.data, dp.ctypes.data,
                                         di.ctypes.data, cap, ctypes.byref(n))
        self._ck(rc)
        return ur[:n.value].copy(), dp[:n.value].copy(), di[:n.value].copy(), rc

    def distinctive_descriptors(self, desc, offsets):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-306) for a batch of map points: point p owns
        desc[offsets[p]:offsets[p+1]].  Returns (best_idx[np] relative to the point's first row, best_median[np])."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        off = np.ascontiguousarray(offsets, np.int32)
        npts = len(off) - 1
        best = np.zeros(max(npts, 0), np.int32); med = np.zeros(max(npts, 0), np.int32)
        self._ck(self._lib.orbx_distinctive_descriptors(self._h, d.ctypes.data, off.ctypes.data, npts, best.ctypes.data, med.ctypes.data))
        return best, med

    def extract_color(self, image, rgb=True):
        """Colour image [H, W, 3 or 4] uint8 -> (keypoints, descriptors); the grey conversion of Tracking::GrabImage*
        (src/Tracking.cc:459-472, `rgb` = Camera.RGB) runs on the device."""
        a = np.ascontiguousarray(image, np.uint8)
        if a.ndim != 3 or a.shape[2] not in (3, 4):
            raise ValueError("expected an [H, W, 3 or 4] uint8 image")
        h, w, c = a.shape
        cap = self.max_keypoints(w, h)
        kps = np.zeros(cap, KEYPOINT_DTYPE); desc = np.zeros((cap, 32), np.uint8)
        n = ctypes.c_int(0)
        self._ck(self._lib.orbx_extract_color(self._h, a.ctypes.data, w, h, a.strides[0], c, int(bool(rgb)), kps.ctypes.data, desc.ctypes.data,
                                              cap, ctypes.byref(n)))
        self._last_batch = 1
        return kps[:n.value], desc[:n.value]
