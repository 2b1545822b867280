"""Python mirror of the reference's ORBVocabulary (include/ORBVocabulary.h: DBoW2::TemplatedVocabulary over FORB) for the
part Frame::ComputeBoW uses (src/Frame.cc:778-785): loadFromTextFile and transform(features, BowVector, FeatureVector,
levelsup).  The descent of every descriptor down the tree runs on the GPU; BowVector / FeatureVector come back as
dict-like ordered structures equal to DBoW2's maps."""
import ctypes

import numpy as np

from . import _lib


class ORBVocabulary:
    def __init__(self, extractor):
        """`extractor` lends its GPU handle (device, stream)."""
        self._ext = extractor
        self._lib = _lib.load_library()
        self._v = ctypes.c_void_p(None)

    def __del__(self):
        if getattr(self, "_v", None) and self._v.value:
            self._lib.orbx_voc_destroy(self._v)
            self._v = ctypes.c_void_p(None)

    def _ck(self, rc):
        return _lib.check(self._lib, self._ext._h, rc)

    def loadFromTextFile(self, filename):
        """TemplatedVocabulary::loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1418)."""
        self.__del__()
        self._ck(self._lib.orbx_voc_load_text(self._ext._h, str(filename).encode(), ctypes.byref(self._v)))
        return True

    def info(self):
        k, L, nodes, words = (ctypes.c_int() for _ in range(4))
        self._ck(self._lib.orbx_voc_info(self._v, ctypes.byref(k), ctypes.byref(L), ctypes.byref(nodes), ctypes.byref(words)))
        return dict(k=k.value, L=L.value, nodes=nodes.value, words=words.value)

    def transform_each(self, descriptors, levelsup=4):
        """Per descriptor: (word id, node id at level L - levelsup, word weight)."""
        d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
        n = len(d)
        word = np.zeros(n, np.int32); node = np.zeros(n, np.int32); weight = np.zeros(n, np.float64)
        self._ck(self._lib.orbx_voc_transform(self._ext._h, self._v, d.ctypes.data, n, int(levelsup), word.ctypes.data, node.ctypes.data,
                                              weight.ctypes.data))
        return word, node, weight

    def transform(self, descriptors, levelsup=4):
        """transform(features, BowVector, FeatureVector, levelsup) flattened in map order:
        (bow_ids, bow_vals, fv_nodes, fv_off, fv_feats)."""
        word, node, weight = self.transform_each(descriptors, levelsup)
        n = len(word)
        bi = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64)
        fn = np.zeros(max(n, 1), np.int32); fo = np.zeros(n + 1, np.int32); ff = np.zeros(max(n, 1), np.int32)
        nb, nf = ctypes.c_int(0), ctypes.c_int(0)
        self._ck(self._lib.orbx_voc_bow(self._v, n, word.ctypes.data, node.ctypes.data, weight.ctypes.data, bi.ctypes.data, bv.ctypes.data,
                                        ctypes.byref(nb), fn.ctypes.data, fo.ctypes.data, ff.ctypes.data, ctypes.byref(nf)))
        return bi[:nb.value].copy(), bv[:nb.value].copy(), fn[:nf.value].copy(), fo[:nf.value + 1].copy(), ff[:fo[nf.value]].copy()
