"""ctypes bindings of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

  Oracle      : oracle/liborb_oracle.so, our C restatement (orb_oracle.c + cvprim.c)
  RefExtractor: oracle/_ref/liborbref_{canon,asis}.so, the reference's own
                src/ORBextractor.cc compiled against oracle/minicv (built here from
                /root/reference by `make -C oracle ref`; the .so travels to the GPU box)
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
_I, _VP, _F = ctypes.c_int, ctypes.c_void_p, ctypes.c_float


def build(ref=True):
    """Compile the port, and the reference-based libs when /root/reference is present."""
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.exists(os.environ.get("ORBX_REFERENCE", "/root/reference") + "/src/ORBextractor.cc"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _load(path):
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    return ctypes.CDLL(path)


class Oracle:
    """Our scalar C port of ORBextractor / ORBmatcher (canonical octree tie-break)."""
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            path = os.path.join(HERE, "liborb_oracle.so")
            if not os.path.exists(path):
                build(ref=False)
            L = _load(path)
            L.orbo_create.restype = _VP
            L.orbo_create.argtypes = [_I, _F, _I, _I, _I]
            L.orbo_destroy.argtypes = [_VP]
            L.orbo_extract.argtypes = [_VP, _VP, _I, _I, _I, _VP, _VP, _I]
            L.orbo_tables.argtypes = [_VP] * 7
            L.orbo_level_size.argtypes = [_VP, _I, _VP, _VP]
            L.orbo_level_image.restype = _VP
            L.orbo_level_image.argtypes = [_VP, _I]
            L.orbo_level_blurred.restype = _VP
            L.orbo_level_blurred.argtypes = [_VP, _I]
            L.orbo_level_candidates.argtypes = [_VP, _I, _VP]
            L.orbo_level_min_cells.argtypes = [_VP, _I, _VP]
            L.orbo_level_nkeypoints.argtypes = [_VP, _I]
            L.orbo_level_padded.argtypes = [_VP, _I, _VP, _I]
            L.orbo_distribute.argtypes = [_VP, _I, _I, _I, _I, _I, _I, _VP, _I]
            L.orbo_hamming256.argtypes = [_VP, _VP]
            L.orbo_match.argtypes = [_VP, _I, _VP, _I, _I, _F, _VP, _VP, _VP, _VP]
            L.orbo_match_mt.argtypes = [_VP, _I, _VP, _I, _I, _F, _VP, _VP, _VP, _VP, _I]
            L.orbo_rotation_filter.argtypes = [_I, _VP, _VP, _VP, _VP, _VP, _VP]
            L.orbo_extract_many.restype = ctypes.c_double
            L.orbo_extract_many.argtypes = [_I, _F, _I, _I, _I, _VP, _I, _I, _I, _I, _VP]
            cls._lib = L
        return cls._lib

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        self.L = self.lib()
        self.params = (int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        self.h = self.L.orbo_create(*self.params)
        if not self.h:
            raise ValueError("orbo_create failed")
        self.nlevels, self.nfeatures = int(nlevels), int(nfeatures)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbo_destroy(self.h)
            self.h = None

    def tables(self):
        L = self.nlevels
        t = [np.zeros(L, np.float32) for _ in range(4)] + [np.zeros(L, np.int32), np.zeros(16, np.int32)]
        self.L.orbo_tables(self.h, *[a.ctypes.data for a in t])
        return dict(scale=t[0], inv_scale=t[1], sigma2=t[2], inv_sigma2=t[3], nfeat=t[4], umax=t[5])

    def __call__(self, image):
        a = np.ascontiguousarray(image, np.uint8)
        cap = self.nfeatures + 8 * self.nlevels + 64
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = self.L.orbo_extract(self.h, a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], kps.ctypes.data, desc.ctypes.data, cap)
        assert n >= 0
        return kps[:n].copy(), desc[:n].copy()

    def level_size(self, level):
        w, h = _I(), _I()
        self.L.orbo_level_size(self.h, level, ctypes.byref(w), ctypes.byref(h))
        return w.value, h.value

    def _img(self, ptr, level):
        w, h = self.level_size(level)
        if not ptr:
            return None
        return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint8)), (h, w)).copy()

    def level_image(self, level):
        return self._img(self.L.orbo_level_image(self.h, level), level)

    def level_blurred(self, level):
        return self._img(self.L.orbo_level_blurred(self.h, level), level)

    def level_padded(self, level):
        w, h = self.level_size(level)
        out = np.zeros((h + 38, w + 38), np.uint8)
        self.L.orbo_level_padded(self.h, level, out.ctypes.data, out.strides[0])
        return out

    def level_candidates(self, level):
        p = _VP()
        n = self.L.orbo_level_candidates(self.h, level, ctypes.byref(p))
        if n == 0:
            return np.zeros((0, 3), np.int32)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_int32)), (n, 3)).copy()

    def level_min_cells(self, level):
        c = _I()
        m = self.L.orbo_level_min_cells(self.h, level, ctypes.byref(c))
        return m, c.value

    def level_nkeypoints(self, level):
        return self.L.orbo_level_nkeypoints(self.h, level)

    @classmethod
    def hamming256(cls, a, b):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return cls.lib().orbo_hamming256(a.ctypes.data, b.ctypes.data)

    @classmethod
    def match(cls, A, B, th, ratio, threads=1):
        A = np.ascontiguousarray(A, np.uint8).reshape(-1, 32); B = np.ascontiguousarray(B, np.uint8).reshape(-1, 32)
        nA, nB = len(A), len(B)
        idx = np.zeros(nA, np.int32); d1 = np.zeros(nA, np.int32); d2 = np.zeros(nA, np.int32); acc = np.zeros(nA, np.uint8)
        if threads > 1:
            cls.lib().orbo_match_mt(A.ctypes.data, nA, B.ctypes.data, nB, int(th), float(ratio), idx.ctypes.data, d1.ctypes.data,
                                    d2.ctypes.data, acc.ctypes.data, int(threads))
        else:
            cls.lib().orbo_match(A.ctypes.data, nA, B.ctypes.data, nB, int(th), float(ratio), idx.ctypes.data, d1.ctypes.data,
                                 d2.ctypes.data, acc.ctypes.data)
        return idx, d1, d2, acc.astype(bool)

    @classmethod
    def distinctive_descriptors(cls, desc, offsets):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:272-301) for a batch of map points: point p owns
        desc[offsets[p]:offsets[p+1]].  Returns (best_idx[np] relative to the point's first row, best_median[np])."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32); off = np.asarray(offsets, np.int64)
        L = cls.lib()
        L.orbo_distinctive_descriptor.argtypes = [_VP, _I, _VP]
        best = np.zeros(len(off) - 1, np.int32); med = np.zeros(len(off) - 1, np.int32)
        for p in range(len(off) - 1):
            m = ctypes.c_int32(0)
            n = int(off[p + 1] - off[p])
            best[p] = L.orbo_distinctive_descriptor(d[off[p]:].ctypes.data, n, ctypes.byref(m)) if n > 0 else -1
            med[p] = m.value if n > 0 else -1
        return best, med

    @classmethod
    def stereo_matches(cls, kpsL, descL, kpsR, descR, scale, inv_scale, pyrL, pyrR, bf):
        """Frame::ComputeStereoMatches (src/Frame.cc:849-1038).  pyrL / pyrR: lists of un-padded level images.
        Returns dict(u_right, depth, desc_index, best_dist, best_idx, sad, kept)."""
        a = _stereo_args(kpsL, descL, kpsR, descR, scale, inv_scale, pyrL, pyrR)
        nL = a["nL"]
        out = dict(u_right=np.zeros(nL, np.float32), depth=np.zeros(nL, np.float32), desc_index=np.zeros(nL, np.int32),
                   best_dist=np.zeros(nL, np.int32), best_idx=np.zeros(nL, np.int32), sad=np.zeros(nL, np.int32))
        L = cls.lib()
        L.orbo_stereo_matches.argtypes = [_I, _VP, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F] + [_VP] * 6
        out["kept"] = L.orbo_stereo_matches(nL, a["kL"].ctypes.data, a["dL"].ctypes.data, a["nR"], a["kR"].ctypes.data, a["dR"].ctypes.data,
                                            a["nlevels"], a["scale"].ctypes.data, a["inv"].ctypes.data, a["pL"], a["pR"],
                                            a["pitch"].ctypes.data, a["lw"].ctypes.data, a["lh"].ctypes.data, float(bf),
                                            *[out[k].ctypes.data for k in ("u_right", "depth", "desc_index", "best_dist", "best_idx", "sad")])
        return out

    @classmethod
    def rotation_filter(cls, idx, accept, angleA, angleB):
        """mbCheckOrientation: returns (accept_after, hist[30], top3[3])."""
        idx = np.ascontiguousarray(idx, np.int32); acc = np.ascontiguousarray(accept, np.uint8).copy()
        a = np.ascontiguousarray(angleA, np.float32); b = np.ascontiguousarray(angleB, np.float32)
        hist = np.zeros(30, np.int32); top3 = np.zeros(3, np.int32)
        cls.lib().orbo_rotation_filter(len(idx), idx.ctypes.data, acc.ctypes.data, a.ctypes.data, b.ctypes.data, hist.ctypes.data, top3.ctypes.data)
        return acc.astype(bool), hist, top3

    @classmethod
    def extract_many(cls, params, frames, threads):
        """CPU baseline: frames [F,H,W] uint8 contiguous; returns (seconds, total keypoints)."""
        f = np.ascontiguousarray(frames, np.uint8)
        tot = ctypes.c_long()
        s = cls.lib().orbo_extract_many(int(params[0]), float(params[1]), int(params[2]), int(params[3]), int(params[4]),
                                        f.ctypes.data, f.shape[0], f.shape[2], f.shape[1], int(threads), ctypes.byref(tot))
        return s, tot.value


def _stereo_args(kpsL, descL, kpsR, descR, scale, inv_scale, pyrL, pyrR):
    kL = np.ascontiguousarray(kpsL, KEYPOINT_DTYPE); kR = np.ascontiguousarray(kpsR, KEYPOINT_DTYPE)
    dL = np.ascontiguousarray(descL, np.uint8).reshape(-1, 32); dR = np.ascontiguousarray(descR, np.uint8).reshape(-1, 32)
    nlevels = len(pyrL)
    keep = [np.ascontiguousarray(p, np.uint8) for p in list(pyrL) + list(pyrR)]
    for l in range(nlevels):
        assert keep[l].shape == keep[nlevels + l].shape
    ptr = ctypes.c_void_p * nlevels
    return dict(kL=kL, kR=kR, dL=dL, dR=dR, nL=len(kL), nR=len(kR), nlevels=nlevels,
                scale=np.ascontiguousarray(scale, np.float32), inv=np.ascontiguousarray(inv_scale, np.float32),
                pL=ptr(*[p.ctypes.data for p in keep[:nlevels]]), pR=ptr(*[p.ctypes.data for p in keep[nlevels:]]),
                pitch=np.array([p.strides[0] for p in keep[:nlevels]], np.int32), lw=np.array([p.shape[1] for p in keep[:nlevels]], np.int32),
                lh=np.array([p.shape[0] for p in keep[:nlevels]], np.int32), _keep=keep)


def _kps_to_rows(k):
    r = np.zeros((len(k), 7), np.float32)
    for i, name in enumerate(["x", "y", "size", "angle", "response", "octave", "class_id"]):
        r[:, i] = k[name]
    return r


class RefExtractor:
    """The reference's own compiled ORBextractor (variant 'canon' or 'asis')."""
    _libs = {}

    @classmethod
    def available(cls, variant="canon"):
        return os.path.exists(os.path.join(HERE, "_ref", "liborbref_%s.so" % variant))

    @classmethod
    def lib(cls, variant):
        if variant not in cls._libs:
            L = _load(os.path.join(HERE, "_ref", "liborbref_%s.so" % variant))
            L.orbref_create.restype = _VP
            L.orbref_create.argtypes = [_I, _F, _I, _I, _I]
            L.orbref_destroy.argtypes = [_VP]
            L.orbref_extract.argtypes = [_VP, _VP, _I, _I, _I, _VP, _VP, _I]
            L.orbref_pyramid_level.argtypes = [_VP, _I, _VP, _I, _VP, _VP, _I]
            L.orbref_tables.argtypes = [_VP] * 7
            L.orbref_distribute.argtypes = [_VP, _VP, _I, _I, _I, _I, _I, _I, _I, _VP, _I]
            L.orbref_descriptor_distance.argtypes = [_VP, _VP]
            L.orbref_thresholds.argtypes = [_VP, _VP, _VP]
            if hasattr(L, "orbref_rotation_filter"):
                L.orbref_rotation_filter.argtypes = [_I, _VP, _VP, _VP, _VP, _VP, _VP]
            if hasattr(L, "orbref_extract_many"):
                L.orbref_extract_many.restype = ctypes.c_double
                L.orbref_extract_many.argtypes = [_I, _F, _I, _I, _I, _VP, _I, _I, _I, _I, _VP]
            cls._libs[variant] = L
        return cls._libs[variant]

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, variant="canon"):
        self.L = self.lib(variant)
        self.h = self.L.orbref_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST))
        self.nlevels, self.nfeatures = int(nlevels), int(nfeatures)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbref_destroy(self.h)
            self.h = None

    def __call__(self, image):
        a = np.ascontiguousarray(image, np.uint8)
        cap = self.nfeatures + 8 * self.nlevels + 64
        k = np.zeros((cap, 7), np.float32)
        d = np.zeros((cap, 32), np.uint8)
        n = self.L.orbref_extract(self.h, a.ctypes.data, a.shape[1], a.shape[0], a.strides[0], k.ctypes.data, d.ctypes.data, cap)
        assert n >= 0
        kps = np.zeros(n, KEYPOINT_DTYPE)
        for i, name in enumerate(["x", "y", "size", "angle", "response"]):
            kps[name] = k[:n, i]
        kps["octave"] = k[:n, 5].astype(np.int32)
        kps["class_id"] = k[:n, 6].astype(np.int32)
        return kps, d[:n].copy()

    def pyramid_level(self, level, with_border=False):
        w, h = _I(), _I()
        self.L.orbref_pyramid_level(self.h, level, None, 0, ctypes.byref(w), ctypes.byref(h), int(with_border))
        out = np.zeros((h.value, w.value), np.uint8)
        self.L.orbref_pyramid_level(self.h, level, out.ctypes.data, out.strides[0], ctypes.byref(w), ctypes.byref(h), int(with_border))
        return out

    def tables(self):
        L = self.nlevels
        t = [np.zeros(L, np.float32) for _ in range(4)] + [np.zeros(L, np.int32), np.zeros(16, np.int32)]
        self.L.orbref_tables(self.h, *[a.ctypes.data for a in t])
        return dict(scale=t[0], inv_scale=t[1], sigma2=t[2], inv_sigma2=t[3], nfeat=t[4], umax=t[5])

    @classmethod
    def descriptor_distance(cls, a, b, variant="canon"):
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        return cls.lib(variant).orbref_descriptor_distance(a.ctypes.data, b.ctypes.data)

    @classmethod
    def distinctive_descriptor(cls, desc, variant="canon"):
        """The reference's own loop (src/MapPoint.cc:272-301) on one map point's descriptors: (best index, its median)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        L = cls.lib(variant)
        L.orbref_distinctive_descriptor.argtypes = [_VP, _I, _VP]
        m = ctypes.c_int(0)
        return L.orbref_distinctive_descriptor(d.ctypes.data, len(d), ctypes.byref(m)), m.value

    @classmethod
    def stereo_matches(cls, kpsL, descL, kpsR, descR, scale, inv_scale, pyrL, pyrR, bf, variant="canon"):
        """The reference's own compiled Frame::ComputeStereoMatches (excerpt of src/Frame.cc:849-1038)."""
        a = _stereo_args(kpsL, descL, kpsR, descR, scale, inv_scale, pyrL, pyrR)
        assert (a["pitch"] == [p.strides[0] for p in a["_keep"][a["nlevels"]:]]).all()
        nL = a["nL"]
        rL, rR = _kps_to_rows(a["kL"]), _kps_to_rows(a["kR"])
        out = dict(u_right=np.zeros(nL, np.float32), depth=np.zeros(nL, np.float32), desc_index=np.zeros(nL, np.int32))
        L = cls.lib(variant)
        L.orbref_stereo_matches.argtypes = [_I, _VP, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _F, _VP, _VP, _VP]
        out["kept"] = L.orbref_stereo_matches(nL, rL.ctypes.data, a["dL"].ctypes.data, a["nR"], rR.ctypes.data, a["dR"].ctypes.data,
                                              a["nlevels"], a["scale"].ctypes.data, a["inv"].ctypes.data, a["pL"], a["pR"],
                                              a["pitch"].ctypes.data, a["lw"].ctypes.data, a["lh"].ctypes.data, float(bf),
                                              out["u_right"].ctypes.data, out["depth"].ctypes.data, out["desc_index"].ctypes.data)
        return out

    @classmethod
    def rotation_filter(cls, idx, accept, angleA, angleB, variant="canon"):
        """The reference's own ComputeThreeMaxima around the call-site code of SearchByBoW (:610-620, 641-660)."""
        idx = np.ascontiguousarray(idx, np.int32); acc = np.ascontiguousarray(accept, np.uint8).copy()
        a = np.ascontiguousarray(angleA, np.float32); b = np.ascontiguousarray(angleB, np.float32)
        hist = np.zeros(30, np.int32); top3 = np.zeros(3, np.int32)
        cls.lib(variant).orbref_rotation_filter(len(idx), idx.ctypes.data, acc.ctypes.data, a.ctypes.data, b.ctypes.data, hist.ctypes.data, top3.ctypes.data)
        return acc.astype(bool), hist, top3

    @classmethod
    def extract_many(cls, params, frames, threads, variant="asis"):
        f = np.ascontiguousarray(frames, np.uint8)
        tot = ctypes.c_long()
        s = cls.lib(variant).orbref_extract_many(int(params[0]), float(params[1]), int(params[2]), int(params[3]), int(params[4]),
                                                 f.ctypes.data, f.shape[0], f.shape[2], f.shape[1], int(threads), ctypes.byref(tot))
        return s, tot.value


def parse_voc_text(path):
    """Rows of an ORBvoc text file (TemplatedVocabulary.h:1338-1418): header `k L scoring weighting`, then per node
    `parent isLeaf d0 .. d31 weight` in file order (node ids 1, 2, ...)."""
    with open(path) as f:
        k, L, scoring, weighting = (int(v) for v in f.readline().split())
        rows = [ln.split() for ln in f if ln.strip()]
    parent = np.array([int(r[0]) for r in rows], np.int32)
    leaf = np.array([int(r[1]) > 0 for r in rows], np.uint8)
    desc = np.array([[int(v) for v in r[2:34]] for r in rows], np.uint8).reshape(-1, 32)
    weight = np.array([float(r[34]) for r in rows], np.float64)
    return k, L, scoring, weighting, parent, leaf, desc, weight


class _VocBase:
    def transform(self, desc, levelsup=4):
        """Frame::ComputeBoW shape: returns (bow_ids, bow_vals, fv_nodes, fv_off, fv_feats)."""
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        word, node, weight = self.transform_each(d, levelsup)
        return self.bow(word, node, weight)


class OracleVocabulary(_VocBase):
    """The C port of the DBoW2 tree (oracle/orb_oracle.c)."""

    def __init__(self, path):
        L = Oracle.lib()
        k, lv, sc, wt, parent, leaf, desc, weight = parse_voc_text(path)
        L.orbo_voc_create.restype = _VP
        L.orbo_voc_create.argtypes = [_I, _I, _I, _I, _I, _VP, _VP, _VP, _VP]
        L.orbo_voc_destroy.argtypes = [_VP]
        L.orbo_voc_transform_each.argtypes = [_VP, _VP, _I, _I, _VP, _VP, _VP]
        L.orbo_voc_bow.argtypes = [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP]
        self.L, self.k, self.levels, self.nodes, self.words = L, k, lv, len(parent) + 1, int(leaf.sum())
        self.h = L.orbo_voc_create(k, lv, sc, wt, len(parent), parent.ctypes.data, leaf.ctypes.data, desc.ctypes.data, weight.ctypes.data)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbo_voc_destroy(self.h); self.h = None

    def transform_each(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        word = np.zeros(n, np.int32); node = np.zeros(n, np.int32); weight = np.zeros(n, np.float64)
        self.L.orbo_voc_transform_each(self.h, d.ctypes.data, n, int(levelsup), word.ctypes.data, node.ctypes.data, weight.ctypes.data)
        return word, node, weight

    def bow(self, word, node, weight):
        n = len(word)
        bi = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64)
        fn = np.zeros(max(n, 1), np.int32); fo = np.zeros(n + 2, np.int32); ff = np.zeros(max(n, 1), np.int32)
        nf = _I(0)
        nb = self.L.orbo_voc_bow(self.h, n, np.ascontiguousarray(word, np.int32).ctypes.data, np.ascontiguousarray(node, np.int32).ctypes.data,
                                 np.ascontiguousarray(weight, np.float64).ctypes.data, bi.ctypes.data, bv.ctypes.data, fn.ctypes.data, fo.ctypes.data,
                                 ff.ctypes.data, ctypes.byref(nf))
        return bi[:nb].copy(), bv[:nb].copy(), fn[:nf.value].copy(), fo[:nf.value + 1].copy(), ff[:fo[nf.value]].copy()


class RefVocabulary(_VocBase):
    """The reference's vendored DBoW2 (Thirdparty/DBoW2) compiled unmodified: oracle/_ref/liborbref_bow.so."""
    _lib = None

    @classmethod
    def available(cls):
        return os.path.exists(os.path.join(HERE, "_ref", "liborbref_bow.so"))

    def __init__(self, path):
        if RefVocabulary._lib is None:
            L = _load(os.path.join(HERE, "_ref", "liborbref_bow.so"))
            L.orbref_voc_load.restype = _VP
            L.orbref_voc_load.argtypes = [ctypes.c_char_p]
            L.orbref_voc_free.argtypes = [_VP]
            L.orbref_voc_info.argtypes = [_VP] * 5
            L.orbref_voc_transform_each.argtypes = [_VP, _VP, _I, _I, _VP, _VP, _VP]
            L.orbref_voc_transform.argtypes = [_VP, _VP, _I, _I, _VP, _VP, _VP, _VP, _VP, _VP]
            RefVocabulary._lib = L
        self.L = RefVocabulary._lib
        self.h = self.L.orbref_voc_load(path.encode())
        assert self.h, "DBoW2 could not load " + path
        k, lv, nn, nw = _I(), _I(), _I(), _I()
        self.L.orbref_voc_info(self.h, ctypes.byref(k), ctypes.byref(lv), ctypes.byref(nn), ctypes.byref(nw))
        self.k, self.levels, self.nodes, self.words = k.value, lv.value, nn.value, nw.value

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orbref_voc_free(self.h); self.h = None

    def transform_each(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        word = np.zeros(n, np.int32); node = np.zeros(n, np.int32); weight = np.zeros(n, np.float64)
        self.L.orbref_voc_transform_each(self.h, d.ctypes.data, n, int(levelsup), word.ctypes.data, node.ctypes.data, weight.ctypes.data)
        return word, node, weight

    def transform(self, desc, levelsup=4):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(d)
        bi = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64)
        fn = np.zeros(max(n, 1), np.int32); fo = np.zeros(n + 2, np.int32); ff = np.zeros(max(n, 1), np.int32)
        nf = _I(0)
        nb = self.L.orbref_voc_transform(self.h, d.ctypes.data, n, int(levelsup), bi.ctypes.data, bv.ctypes.data, fn.ctypes.data, fo.ctypes.data,
                                         ff.ctypes.data, ctypes.byref(nf))
        return bi[:nb].copy(), bv[:nb].copy(), fn[:nf.value].copy(), fo[:nf.value + 1].copy(), ff[:fo[nf.value]].copy()


def _projection_args(case):
    c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
    return [c["cam"].ctypes.data, c["Tcw_cur"].ctypes.data, c["Tcw_last"].ctypes.data, len(c["world_pos"]), c["world_pos"].ctypes.data,
            c["mp_desc"].ctypes.data, c["valid"].ctypes.data, c["nobs"].ctypes.data, c["last_octave"].ctypes.data, c["last_angle"].ctypes.data,
            len(c["cur_xy"]), c["cur_xy"].ctypes.data, c["cur_octave"].ctypes.data, c["cur_angle"].ctypes.data, c["cur_uright"].ctypes.data,
            c["cur_desc"].ctypes.data, c["scale"].ctypes.data, len(c["scale"]), float(case["th"]), int(case["mono"]), int(case["check_orientation"])], c


_PROJ_ARGTYPES = [_VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _F, _I, _I, _VP]


def search_by_projection_port(case):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, ...) by the C port: (cur_match[nC], nmatches)."""
    L = Oracle.lib()
    L.orbo_search_by_projection.argtypes = _PROJ_ARGTYPES
    args, keep = _projection_args(case)
    out = np.full(len(keep["cur_xy"]), -1, np.int32)
    n = L.orbo_search_by_projection(*args, out.ctypes.data)
    return out, n


def search_by_projection_ref(case, variant="canon"):
    """The reference's own compiled SearchByProjection + Frame grid (excerpts of src/ORBmatcher.cc:1958-2102, src/Frame.cc)."""
    L = RefExtractor.lib(variant)
    L.orbref_search_by_projection.argtypes = _PROJ_ARGTYPES
    args, keep = _projection_args(case)
    out = np.full(len(keep["cur_xy"]), -1, np.int32)
    n = L.orbref_search_by_projection(*args, out.ctypes.data)
    return out, n


_INIT_ARGTYPES = [_VP, _I, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _I, _F, _I, _VP]


def _init_call(fn, case):
    c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
    prev = c["prev_xy"].astype(np.float32).copy()
    m12 = np.full(len(c["xy1"]), -1, np.int32)
    fn.argtypes = _INIT_ARGTYPES
    n = fn(c["cam"].ctypes.data, len(c["xy1"]), c["xy1"].ctypes.data, c["oct1"].ctypes.data, c["ang1"].ctypes.data, c["desc1"].ctypes.data,
           len(c["xy2"]), c["xy2"].ctypes.data, c["oct2"].ctypes.data, c["ang2"].ctypes.data, c["desc2"].ctypes.data,
           prev.ctypes.data, int(case["window"]), float(case["nnratio"]), int(case["check_orientation"]), m12.ctypes.data)
    return m12, prev, n


def search_for_initialization_port(case):
    """ORBmatcher::SearchForInitialization by the C port: (vnMatches12, vbPrevMatched after the call, nmatches)."""
    return _init_call(Oracle.lib().orbo_search_for_initialization, case)


def search_for_initialization_ref(case, variant="canon"):
    """The reference's own compiled SearchForInitialization (excerpt of src/ORBmatcher.cc:780-895)."""
    return _init_call(RefExtractor.lib(variant).orbref_search_for_initialization, case)


_LOCAL_ARGTYPES = [_VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _VP, _VP, _VP, _VP, _VP, _VP, _I, _F, _F, _VP]


def _local_call(fn, case):
    c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
    out = np.full(len(c["xy"]), -1, np.int32)
    fn.argtypes = _LOCAL_ARGTYPES
    n = fn(c["cam"].ctypes.data, len(c["proj"]), c["proj"].ctypes.data, c["view_cos"].ctypes.data, c["level"].ctypes.data, c["mp_desc"].ctypes.data,
           c["valid"].ctypes.data, c["nobs"].ctypes.data, len(c["xy"]), c["xy"].ctypes.data, c["octave"].ctypes.data, c["uright"].ctypes.data,
           c["desc"].ctypes.data, c["feat_obs"].ctypes.data, c["scale"].ctypes.data, len(c["scale"]), float(case["th"]), float(case["nnratio"]),
           out.ctypes.data)
    return out, n


def search_local_points_port(case):
    """ORBmatcher::SearchByProjection(F, vpMapPoints, th) by the C port: (index of the newly assigned map point per feature, nmatches)."""
    return _local_call(Oracle.lib().orbo_search_local_points, case)


def search_local_points_ref(case, variant="canon"):
    """The reference's own compiled function (excerpt of src/ORBmatcher.cc:418-511)."""
    return _local_call(RefExtractor.lib(variant).orbref_search_local_points, case)


_BOWM_ARGTYPES = [_I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP, _F, _I, _VP]


def _bowm_call(fn, case):
    c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
    out = np.full(len(c["f_desc"]), -1, np.int32)
    fn.argtypes = _BOWM_ARGTYPES
    n = fn(len(c["kf_desc"]), c["kf_angle"].ctypes.data, c["kf_desc"].ctypes.data, c["kf_valid"].ctypes.data, len(c["kf_nodes"]), c["kf_nodes"].ctypes.data,
           c["kf_off"].ctypes.data, c["kf_feats"].ctypes.data, len(c["f_desc"]), c["f_angle"].ctypes.data, c["f_desc"].ctypes.data, len(c["f_nodes"]),
           c["f_nodes"].ctypes.data, c["f_off"].ctypes.data, c["f_feats"].ctypes.data, float(case["nnratio"]), int(case["check_orientation"]), out.ctypes.data)
    return out, n


def search_by_bow_port(case):
    """ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...) by the C port: (key-frame feature matched to every frame feature, nmatches)."""
    return _bowm_call(Oracle.lib().orbo_search_by_bow, case)


def search_by_bow_ref(case, variant="canon"):
    """The reference's own compiled function (excerpt of src/ORBmatcher.cc:532-663)."""
    return _bowm_call(RefExtractor.lib(variant).orbref_search_by_bow, case)


_BOWK_ARGTYPES = [_I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _I, _VP, _VP, _VP, _F, _I, _VP]


def _bowk_call(fn, case):
    c = {k: np.ascontiguousarray(v) for k, v in case.items() if isinstance(v, np.ndarray)}
    out = np.full(len(c["kf_desc"]), -1, np.int32)
    fn.argtypes = _BOWK_ARGTYPES
    n = fn(len(c["kf_desc"]), c["kf_angle"].ctypes.data, c["kf_desc"].ctypes.data, c["kf_valid"].ctypes.data, len(c["kf_nodes"]), c["kf_nodes"].ctypes.data,
           c["kf_off"].ctypes.data, c["kf_feats"].ctypes.data, len(c["f_desc"]), c["f_angle"].ctypes.data, c["f_desc"].ctypes.data, c["f_valid"].ctypes.data,
           len(c["f_nodes"]), c["f_nodes"].ctypes.data, c["f_off"].ctypes.data, c["f_feats"].ctypes.data, float(case["nnratio"]),
           int(case["check_orientation"]), out.ctypes.data)
    return out, n


def search_by_bow_kf_port(case):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) by the C port (key frame 1 = the case's kf_*, key frame 2 = its f_* + f_valid)."""
    return _bowk_call(Oracle.lib().orbo_search_by_bow_kf, case)


def search_by_bow_kf_ref(case, variant="canon"):
    """The reference's own compiled function (excerpt of src/ORBmatcher.cc:897-1030)."""
    return _bowk_call(RefExtractor.lib(variant).orbref_search_by_bow_kf, case)
