"""Python mirror of the frame-sharded dispatcher of include/orbx.h (orbx_pool_*): one native worker thread and `depth`
extractor handles per GPU; a submit only queues the shards, the worker threads do the CUDA launches, results come back as
zero-copy views of pinned host memory per shard (the reference's threading contract is src/Frame.cc:96-99)."""
import ctypes

import numpy as np

from . import _lib
from .extractor import COMPACT_KEYPOINT_DTYPE, KEYPOINT_DTYPE
from .sharding import shard_bounds


class ExtractorPool:
    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, devices=None, depth=6, max_width=0, max_height=0, max_batch=0,
                 compact_keypoints=False):
        self._lib = _lib.load_library()
        self._devs = (ctypes.c_int32 * len(devices))(*devices) if devices else None
        cfg = _lib.OrbxPoolConfig(_lib.OrbxConfig(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST),
                                                  int(max_width), int(max_height), int(max_batch), -1),
                                  len(devices) if devices else 0, self._devs, int(depth), int(bool(compact_keypoints)))
        p = ctypes.c_void_p()
        rc = self._lib.orbx_pool_create(ctypes.byref(cfg), ctypes.byref(p))
        if rc != 0:
            raise _lib.OrbxError(rc, (self._lib.orbx_pool_last_error(None) or b"").decode())
        self._p = p
        self.nshards = self._lib.orbx_pool_devices(p)
        self.depth = self._lib.orbx_pool_depth(p)
        self._keep = {}

    def close(self):
        if getattr(self, "_p", None):
            self._lib.orbx_pool_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc < 0:
            raise _lib.OrbxError(int(rc), (self._lib.orbx_pool_last_error(self._p) or b"").decode())
        return rc

    def set_option(self, option, value):
        """ORBX_OPT_* on every handle (no ticket may be outstanding)."""
        self._ck(self._lib.orbx_pool_set_option(self._p, int(option), int(value)))

    def expand_keypoints(self, ckps):
        c = np.ascontiguousarray(ckps, COMPACT_KEYPOINT_DTYPE)
        out = np.zeros(len(c), KEYPOINT_DTYPE)
        self._lib.orbx_expand_keypoints(self.handle_ptr(0, 0), c.ctypes.data, len(c), out.ctypes.data)
        return out

    def shard_range(self, nframes, shard):
        first, count = ctypes.c_int(), ctypes.c_int()
        self._lib.orbx_pool_shard_range(int(nframes), self.nshards, int(shard), ctypes.byref(first), ctypes.byref(count))
        assert (first.value, first.value + count.value) == shard_bounds(nframes, shard, self.nshards)
        return first.value, count.value

    def launch_count(self):
        """Kernel launches issued by all handles of the pool since creation."""
        return sum(int(self._lib.orbx_launch_count(self._lib.orbx_pool_handle(self._p, g, k)))
                   for g in range(self.nshards) for k in range(self.depth))

    def handle_ptr(self, shard, slot):
        return self._lib.orbx_pool_handle(self._p, int(shard), int(slot))

    def submit_host(self, frames):
        """frames: list of HxW uint8 arrays of one shape and row stride (or one [F,H,W] array).  Returns a ticket."""
        arrs = [np.asarray(f) for f in frames]
        h, w = arrs[0].shape
        F = len(arrs)
        ptrs = (ctypes.c_void_p * F)(*[a.ctypes.data for a in arrs])
        t = self._ck(self._lib.orbx_pool_submit_host(self._p, ptrs, F, w, h, arrs[0].strides[0]))
        self._keep[t] = (arrs, ptrs)
        return t

    def submit_device(self, ptrs, nframes, width, height, stride_bytes, frame_stride_bytes):
        """ptrs[g], nframes[g]: device-resident frames of shard g (on that shard's GPU).  Returns a ticket."""
        G = self.nshards
        assert len(ptrs) == G and len(nframes) == G
        pa = (ctypes.c_void_p * G)(*[int(p) for p in ptrs])
        na = (ctypes.c_int32 * G)(*[int(n) for n in nframes])
        return self._ck(self._lib.orbx_pool_submit_device(self._p, pa, na, int(width), int(height), int(stride_bytes), int(frame_stride_bytes)))

    def collect(self, ticket):
        """Waits for the ticket; returns a list of shards: (first_frame, kps [F,cap], desc [F,cap,32], n [F]) as views of the
        pool's pinned buffers (valid until the submit that returns ticket + depth); kps are COMPACT_KEYPOINT_DTYPE records when the
        pool delivers compact keypoints (expand_keypoints rebuilds cv::KeyPoint records)."""
        res = (_lib.OrbxShardResult * self.nshards)()
        self._ck(self._lib.orbx_pool_collect(self._p, int(ticket), res))
        self._keep.pop(ticket, None)
        out = []
        for r in res:
            F, c = r.nframes, r.cap_per_frame
            if F == 0:
                out.append((r.first_frame, np.zeros((0, 0), KEYPOINT_DTYPE), np.zeros((0, 0, 32), np.uint8), np.zeros(0, np.int32)))
                continue
            if r.ckps:
                kb = (ctypes.c_uint8 * (F * c * 12)).from_address(r.ckps)
                kps = np.frombuffer(kb, dtype=COMPACT_KEYPOINT_DTYPE).reshape(F, c)
            else:
                kb = (ctypes.c_uint8 * (F * c * 28)).from_address(r.kps)
                kps = np.frombuffer(kb, dtype=KEYPOINT_DTYPE).reshape(F, c)
            db = (ctypes.c_uint8 * (F * c * 32)).from_address(r.desc)
            nb = (ctypes.c_int32 * F).from_address(r.n)
            out.append((r.first_frame, kps, np.frombuffer(db, dtype=np.uint8).reshape(F, c, 32), np.frombuffer(nb, dtype=np.int32)))
        return out
