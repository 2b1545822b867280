#!/usr/bin/env python3
"""CPU prototype of the GPU octree formulation (sorted path codes + node ranges),
checked against oracle/orb_oracle.c's sequential restatement of DistributeOctTree
(src/ORBextractor.cc:539-763).  Development aid for kernels.cu::k_octree; the same
algorithm, step for step, as the CUDA kernel (level-synchronous sweeps, careful
phase with a (count, position)-descending visit order and a prefix-sum cutoff)."""
import ctypes, math, sys, os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = ctypes.CDLL(os.path.join(ROOT, "oracle", "liborb_oracle.so"))

def f32(x): return np.float32(x)

def geometry(w, h):
    minX = minY = 16; maxX = w - 16; maxY = h - 16
    ow, oh = maxX - minX, maxY - minY
    nIni = int(np.round(f32(ow) / f32(oh)))          # round half away: values never at .5 exactly matter little here
    nIni = int(math.floor(float(f32(ow) / f32(oh)) + 0.5))
    hX = f32(ow) / f32(nIni)
    ul = [int(hX * f32(i)) for i in range(nIni)]
    br = [int(hX * f32(i + 1)) for i in range(nIni)]
    maxdim = max([oh] + [b - a for a, b in zip(ul, br)])
    D = 0
    while (1 << D) < maxdim: D += 1
    fw, fh = f32(ow), f32(oh)
    nCols = int(fw / f32(30)); nRows = int(fh / f32(30))
    wCell = int(math.ceil(float(fw / f32(nCols)))); hCell = int(math.ceil(float(fh / f32(nRows))))
    return dict(minX=minX, maxX=maxX, minY=minY, maxY=maxY, nIni=nIni, hX=hX, ul=ul, br=br, oh=oh, D=D,
                nCols=nCols, wCell=wCell, hCell=hCell)

def path_key(x, y, G):
    r = int(f32(x) / G["hX"])
    ulx, brx, uly, bry = G["ul"][r], G["br"][r], 0, G["oh"]
    code = 0
    for _ in range(G["D"]):
        midx = ulx + ((brx - ulx + 1) >> 1); midy = uly + ((bry - uly + 1) >> 1)
        right = x >= midx; down = y >= midy
        if right: ulx = midx
        else: brx = midx
        if down: uly = midy
        else: bry = midy
        code = (code << 2) | (int(down) << 1) | int(right)
    return (r << (2 * G["D"])) | code

def distribute_ranges(cands, G, N):
    """cands: list of (x, y, score) relative coords.  Returns selected list in reference order."""
    D = G["D"]
    keyed = sorted(((path_key(x, y, G), (x, y, s)) for (x, y, s) in cands))
    keys = [k for k, _ in keyed]; pay = [p for _, p in keyed]
    n = len(keys)
    import bisect
    def lower(lo, hi, shift, c):       # first i in [lo,hi) with ((key>>shift)&3) >= c
        a, b = lo, hi
        while a < b:
            m = (a + b) // 2
            if ((keys[m] >> shift) & 3) >= c: b = m
            else: a = m + 1
        return a
    # roots, reversed list order
    A = []
    for r in reversed(range(G["nIni"])):
        lo = bisect.bisect_left(keys, r << (2 * D)); hi = bisect.bisect_left(keys, (r + 1) << (2 * D))
        if hi > lo: A.append([lo, hi, 0])
    def children(node):
        lo, hi, d = node
        shift = 2 * (D - 1 - d)
        b = [lo, lower(lo, hi, shift, 1), lower(lo, hi, shift, 2), lower(lo, hi, shift, 3), hi]
        return [[b[c], b[c + 1], d + 1] for c in range(4) if b[c + 1] > b[c]]
    size = len(A)
    mode = "sweep"
    while True:
        prev = size
        E = [i for i, nd in enumerate(A) if nd[1] - nd[0] > 1]
        if mode == "sweep":
            visit = list(reversed(E)); m = len(visit)
            kids = [children(A[p]) for p in visit]
        else:
            visit = sorted(E, key=lambda p: (A[p][1] - A[p][0], p), reverse=True)
            kids = [children(A[p]) for p in visit]
            m = len(visit); acc = 0
            for r, k in enumerate(kids):
                acc += len(k) - 1
                if size + acc >= N: m = r + 1; break
        dead = set(visit[:m])
        added = [c for k in kids[:m] for c in k]
        size = size - m + len(added)
        A = [nd for i, nd in enumerate(A) if i not in dead] + added
        assert len(A) == size
        nE = sum(1 for nd in A if nd[1] - nd[0] > 1)
        if size >= N or size == prev: break
        if mode == "sweep" and size + 3 * nE > N: mode = "careful"
    out = []
    nC, wC, hC = G["nCols"], G["wCell"], G["hCell"]
    for nd in reversed(A):
        best = None
        for i in range(nd[0], nd[1]):
            x, y, s = pay[i]
            order = ((((y - 3) // hC) * nC + (x - 3) // wC) << 24) | (y << 12) | x
            k = (s, -order)
            if best is None or k > best[0]: best = (k, pay[i])
        out.append(best[1])
    return out

if __name__ == "__main__":
    import cv2
    sys.path.insert(0, ROOT)
    O.orbo_create.restype = ctypes.c_void_p
    O.orbo_create.argtypes = [ctypes.c_int, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    O.orbo_extract.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2 + [ctypes.c_int]
    O.orbo_level_candidates.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    O.orbo_level_size.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    O.orbo_distribute.argtypes = [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 5 + [ctypes.c_void_p, ctypes.c_int]
    def valnoise(seed, h, w):
        rng = np.random.default_rng(seed)
        rows, cols = -(-h // 12) + 2, -(-w // 12) + 2
        coarse = rng.integers(0, 256, (rows, cols), dtype=np.uint8)
        up = cv2.resize(coarse, (12 * cols, 12 * rows), interpolation=cv2.INTER_CUBIC)[12:12 + h, 12:12 + w]
        nz = rng.integers(-4, 5, (h, w), dtype=np.int16)
        return np.clip(up.astype(np.int16) + nz, 0, 255).astype(np.uint8)
    frames = [("syn%d" % s, valnoise(s, 375, 1242)) for s in range(3)]
    frames.append(("noise", np.random.default_rng(5).integers(0, 256, (375, 1242), dtype=np.uint8)))
    frames.append(("tall", valnoise(9, 500, 640)))
    if os.path.exists("/root/reference/kitti_sample/image/000000.png"):
        frames.append(("kitti0", cv2.cvtColor(cv2.imread("/root/reference/kitti_sample/image/000000.png", cv2.IMREAD_UNCHANGED), cv2.COLOR_RGB2GRAY)))
    bad = 0
    for name, g in frames:
        for nfeat in (2000, 4000, 300):
            e = O.orbo_create(nfeat, 1.2, 8, 20, 7)
            nf = (ctypes.c_int * 8)()
            O.orbo_tables(ctypes.c_void_p(e), None, None, None, None, nf, None)
            O.orbo_extract(e, g.ctypes.data, g.shape[1], g.shape[0], g.strides[0], None, None, 0)
            for l in range(8):
                p = ctypes.c_void_p(); w = ctypes.c_int(); h = ctypes.c_int()
                n = O.orbo_level_candidates(e, l, ctypes.byref(p))
                O.orbo_level_size(e, l, ctypes.byref(w), ctypes.byref(h))
                c = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_int32)), (n, 3)).copy()
                G = geometry(w.value, h.value)
                out = np.zeros((n + 8, 3), np.int32)
                m = O.orbo_distribute(c.ctypes.data, n, G["minX"], G["maxX"], G["minY"], G["maxY"], nf[l], out.ctypes.data, n + 8)
                ref = [tuple(r) for r in out[:m].tolist()]
                got = distribute_ranges([tuple(r) for r in c.tolist()], G, nf[l])
                ok = ref == got
                bad += not ok
                if not ok: print("MISMATCH", name, nfeat, l, n, m, len(got))
            print(name, nfeat, "done")
    print("bad", bad)
