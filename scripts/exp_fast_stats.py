"""Measurement aid (CPU only): how selective are FAST pre-tests on the benchmark's synthetic frames?

For every pyramid level of value-noise frames (the pool bench.py uses) the exact FAST-9/16 score map is compared with
  p4   the compass pre-test on 4 opposing ring pairs: every pair has a pixel beyond the threshold (what cv::FAST tests first)
  p8   the same on all 8 opposing pairs
  x8   p8 AND one opposing pair has BOTH pixels beyond the threshold in the same direction (every 9-arc contains such a pair):
       the tightest cheap necessary condition (26 packed min / max for both polarities, against 72 for the exact score)
at iniThFAST = 20 and minThFAST = 7: fraction of pixels, of pixel PAIRS (the unit a packed 16x2 lane scores) and of 64-pixel
warp rows that pass.  Decides whether an "iniThFAST first, exact score for the survivors only" kernel can pay.
usage: python scripts/exp_fast_stats.py > profiles/r2_fast_pretest_rates.txt"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import Oracle
from multimot_track_b200.synth import value_noise_frame

OFF = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def score_map(I):
    I = I.astype(np.int16); H, W = I.shape
    c = I[3:H - 3, 3:W - 3]
    r = np.stack([I[3 + dy:H - 3 + dy, 3 + dx:W - 3 + dx] for dx, dy in OFF])
    d = c[None] - r
    d2 = np.concatenate([d, d[:8]], 0)
    best = np.full(c.shape, -999, np.int16)
    for k in range(16):
        arc = d2[k:k + 9]
        best = np.maximum(best, np.maximum(arc.min(0), -arc.max(0)))
    return best - 1, r, c


def seg(m, n):
    w = m.shape[1] // n * n
    return m[:, :w].reshape(m.shape[0], -1, n).any(2).mean()


print("# scripts/exp_fast_stats.py: value-noise frames 1242x375 (seeds 0, 7), 8 levels x1.2; fractions that pass")
print("# level  share-of-pixels | threshold | exact corners: pixel pair row64 | p4: pixel pair | p8: pixel pair row64 | x8: pixel pair row64")
tot = {}
for seed in (0, 7):
    img = value_noise_frame(seed, 375, 1242)
    o = Oracle(2000, 1.2, 8, 20, 7); o(img)
    px = [o.level_image(l).size for l in range(8)]
    for l in range(8):
        S, r, c = score_map(o.level_image(l))
        for t in (20, 7):
            M = [np.maximum(r[k], r[k + 8]) for k in range(8)]; m = [np.minimum(r[k], r[k + 8]) for k in range(8)]
            A, B = np.minimum.reduce(M), np.maximum.reduce(m)
            A4, B4 = np.minimum.reduce(M[0::2]), np.maximum.reduce(m[0::2])
            p8 = (A > c + t) | (B < c - t)
            p4 = (A4 > c + t) | (B4 < c - t)
            x8 = (np.minimum(A, B) > c + t) | (np.maximum(A, B) < c - t)
            tr = S >= t
            assert not (tr & ~x8).any() and not (x8 & ~p8).any()                  # necessary conditions, nested
            row = (tr.mean(), seg(tr, 2), seg(tr, 64), p4.mean(), seg(p4, 2), p8.mean(), seg(p8, 2), seg(p8, 64), x8.mean(), seg(x8, 2), seg(x8, 64))
            tot.setdefault((l, t), []).append((px[l] / sum(px),) + row)
for (l, t), rows in sorted(tot.items()):
    v = np.mean(rows, 0)
    print("L%d  %5.3f | th %2d | exact %.4f %.4f %.3f | p4 %.4f %.4f | p8 %.4f %.4f %.3f | x8 %.4f %.4f %.3f" % ((l, v[0], t) + tuple(v[1:])))
for t in (20, 7):
    w = np.array([np.mean(tot[(l, t)], 0) for l in range(8)])
    print("all levels, pixel-weighted, th %2d: exact pairs %.3f, p8 pairs %.3f, x8 pairs %.3f" % (t, (w[:, 0] * w[:, 2]).sum(), (w[:, 0] * w[:, 7]).sum(), (w[:, 0] * w[:, 10]).sum()))
