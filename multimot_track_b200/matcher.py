"""Python mirror of the hot core of ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-109):
`DescriptorDistance` (src/ORBmatcher.cc:2279-2295), the thresholds (:41-43) and the
best / second-best / ratio scan every SearchBy* shares (:574-605), as an all-pairs
brute-force match on the GPU (BASELINE.json config 4)."""
import numpy as np

from . import _lib


class ORBmatcher:
    TH_LOW = 50
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio=0.6, checkOri=True, extractor=None):
        """nnratio / checkOri as in ORBmatcher::ORBmatcher (src/ORBmatcher.cc:50).  `extractor`
        lends its GPU handle (stream + scratch); a private one is created otherwise."""
        from .extractor import ORBextractor
        self.mfNNratio = float(np.float32(nnratio))
        self.mbCheckOrientation = bool(checkOri)
        self._ext = extractor if extractor is not None else ORBextractor(1000, 1.2, 1, 20, 7)

    @staticmethod
    def DescriptorDistance(a, b):
        """256-bit Hamming distance of two 32-byte descriptors (host scalar, like the static member)."""
        lib = _lib.load_library()
        a = np.frombuffer(a, np.uint8) if isinstance(a, (bytes, bytearray)) else np.ascontiguousarray(a, np.uint8).reshape(-1)
        b = np.frombuffer(b, np.uint8) if isinstance(b, (bytes, bytearray)) else np.ascontiguousarray(b, np.uint8).reshape(-1)
        if a.size != 32 or b.size != 32:
            raise ValueError("descriptors are 32 bytes")
        return int(lib.orbx_hamming256(a.ctypes.data, b.ctypes.data))

    def match(self, descA, descB, th=None, ratio=None):
        """All-pairs scan.  Returns idx[nA] (-1 if nB == 0), best[nA], second[nA], accept[nA] where
        accept = best <= th and float(best) < ratio*float(second)."""
        th = self.TH_LOW if th is None else th
        ratio = self.mfNNratio if ratio is None else ratio
        return self._ext.match(descA, descB, th, ratio)

    def match_oriented(self, descA, anglesA, descB, anglesB, th=None, ratio=None):
        """match() followed, when mbCheckOrientation is set, by the rotation-consistency pruning every SearchBy*
        applies to its accepted matches (src/ORBmatcher.cc:610-620, 641-660; ComputeThreeMaxima :2233-2274).
        Returns idx, best, second, accept (after pruning), hist[30], top3[3]."""
        idx, d1, d2, acc = self.match(descA, descB, th, ratio)
        if not self.mbCheckOrientation:
            return idx, d1, d2, acc, np.zeros(30, np.int32), np.full(3, -1, np.int32)
        acc, hist, top3 = self._ext.rotation_filter(idx, acc, anglesA, anglesB)
        return idx, d1, d2, acc, hist, top3

    def SearchByProjection(self, case):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:1958-2102).  `case` is a dict with
        the arrays of include/orbx.h's orbx_search_by_projection (see multimot_track_b200.synth.projection_case): cam (fx fy cx cy
        mbf mb mnMinX mnMaxX mnMinY mnMaxY), Tcw_cur, Tcw_last, world_pos, mp_desc, valid, nobs, last_octave, last_angle, cur_xy,
        cur_octave, cur_angle, cur_uright, cur_desc, th, mono, check_orientation.  Returns (cur_match[n_cur], nmatches)."""
        import ctypes
        c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
        setup = np.concatenate([c["cam"].astype(np.float32), c["Tcw_cur"].astype(np.float32).reshape(-1), c["Tcw_last"].astype(np.float32).reshape(-1)])
        setup = np.ascontiguousarray(setup, np.float32)
        n_cur = len(c["cur_xy"])
        out = np.full(n_cur, -1, np.int32)
        ext = self._ext
        rc = ext._lib.orbx_search_by_projection(ext._h, setup.ctypes.data, len(c["world_pos"]), c["world_pos"].ctypes.data, c["mp_desc"].ctypes.data,
                                                c["valid"].ctypes.data, c["nobs"].ctypes.data, c["last_octave"].ctypes.data, c["last_angle"].ctypes.data,
                                                n_cur, c["cur_xy"].ctypes.data, c["cur_octave"].ctypes.data, c["cur_angle"].ctypes.data,
                                                c["cur_uright"].ctypes.data, c["cur_desc"].ctypes.data, float(case["th"]), int(case["mono"]),
                                                int(self.mbCheckOrientation if "check_orientation" not in case else case["check_orientation"]), out.ctypes.data)
        ext._ck(rc)
        return out, rc

    def SearchForInitialization(self, case):
        """ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) (src/ORBmatcher.cc:780-895).  `case`
        (multimot_track_b200.synth.initialization_case): cam (.. mnMinX mnMaxX mnMinY mnMaxY at [6:10]), oct1, ang1, desc1, xy2, oct2,
        ang2, desc2, prev_xy (vbPrevMatched), window, nnratio, check_orientation.
        Returns (vnMatches12[n1], vbPrevMatched after the call, nmatches)."""
        c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
        ext = self._ext
        n1, n2 = len(c["oct1"]), len(c["xy2"])
        prev = np.ascontiguousarray(c["prev_xy"], np.float32).copy()
        m12 = np.full(n1, -1, np.int32)
        b = [float(x) for x in c["cam"][6:10]]
        rc = ext._lib.orbx_search_for_initialization(ext._h, b[0], b[1], b[2], b[3], n1, c["oct1"].ctypes.data, c["ang1"].ctypes.data, c["desc1"].ctypes.data,
                                                     n2, c["xy2"].ctypes.data, c["oct2"].ctypes.data, c["ang2"].ctypes.data, c["desc2"].ctypes.data,
                                                     prev.ctypes.data, int(case["window"]), float(case.get("nnratio", self.mfNNratio)),
                                                     int(self.mbCheckOrientation if "check_orientation" not in case else case["check_orientation"]), m12.ctypes.data)
        ext._ck(rc)
        return m12, prev, rc

    def SearchLocalPoints(self, case):
        """ORBmatcher::SearchByProjection(F, vpMapPoints, th) (src/ORBmatcher.cc:418-502), the call of Tracking::SearchLocalPoints.
        `case` (multimot_track_b200.synth.local_points_case): cam (image bounds at [6:10]), proj, view_cos, level, mp_desc, valid, nobs,
        xy, octave, uright, desc, feat_obs, th, nnratio.  Returns (index of the newly assigned map point per feature, nmatches)."""
        c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
        ext = self._ext
        nf = len(c["xy"])
        out = np.full(nf, -1, np.int32)
        b = [float(x) for x in c["cam"][6:10]]
        rc = ext._lib.orbx_search_local_points(ext._h, b[0], b[1], b[2], b[3], len(c["proj"]), c["proj"].ctypes.data, c["view_cos"].ctypes.data,
                                               c["level"].ctypes.data, c["mp_desc"].ctypes.data, c["valid"].ctypes.data, c["nobs"].ctypes.data,
                                               nf, c["xy"].ctypes.data, c["octave"].ctypes.data, c["uright"].ctypes.data, c["desc"].ctypes.data,
                                               c["feat_obs"].ctypes.data, float(case["th"]), float(case.get("nnratio", self.mfNNratio)), out.ctypes.data)
        ext._ck(rc)
        return out, rc

    def SearchByBoW(self, case):
        """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches, ..) (src/ORBmatcher.cc:532-663).  `case`
        (multimot_track_b200.synth.bow_match_case): kf_angle, kf_desc, kf_valid, kf_nodes / kf_off / kf_feats, f_angle, f_desc,
        f_nodes / f_off / f_feats (feature vectors flattened in map order), nnratio, check_orientation.
        Returns (key-frame feature matched to every frame feature, nmatches)."""
        c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
        ext = self._ext
        nf = len(c["f_desc"])
        out = np.full(nf, -1, np.int32)
        rc = ext._lib.orbx_search_by_bow(ext._h, len(c["kf_desc"]), c["kf_angle"].ctypes.data, c["kf_desc"].ctypes.data, c["kf_valid"].ctypes.data,
                                         len(c["kf_nodes"]), c["kf_nodes"].ctypes.data, c["kf_off"].ctypes.data, c["kf_feats"].ctypes.data,
                                         nf, c["f_angle"].ctypes.data, c["f_desc"].ctypes.data, len(c["f_nodes"]), c["f_nodes"].ctypes.data,
                                         c["f_off"].ctypes.data, c["f_feats"].ctypes.data, float(case.get("nnratio", self.mfNNratio)),
                                         int(self.mbCheckOrientation if "check_orientation" not in case else case["check_orientation"]), out.ctypes.data)
        ext._ck(rc)
        return out, rc

    def SearchByBoWKeyFrames(self, case):
        """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (src/ORBmatcher.cc:897-1030).  `case` as for SearchByBoW with key frame 1 =
        kf_* and key frame 2 = f_* plus f_valid.  Returns (feature of key frame 2 matched to every feature of key frame 1, nmatches)."""
        c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
        ext = self._ext
        n1 = len(c["kf_desc"])
        out = np.full(n1, -1, np.int32)
        rc = ext._lib.orbx_search_by_bow_keyframes(ext._h, n1, c["kf_angle"].ctypes.data, c["kf_desc"].ctypes.data, c["kf_valid"].ctypes.data,
                                                   len(c["kf_nodes"]), c["kf_nodes"].ctypes.data, c["kf_off"].ctypes.data, c["kf_feats"].ctypes.data,
                                                   len(c["f_desc"]), c["f_angle"].ctypes.data, c["f_desc"].ctypes.data, c["f_valid"].ctypes.data,
                                                   len(c["f_nodes"]), c["f_nodes"].ctypes.data, c["f_off"].ctypes.data, c["f_feats"].ctypes.data,
                                                   float(case.get("nnratio", self.mfNNratio)),
                                                   int(self.mbCheckOrientation if "check_orientation" not in case else case["check_orientation"]), out.ctypes.data)
        ext._ck(rc)
        return out, rc
