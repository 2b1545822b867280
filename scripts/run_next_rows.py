"""Measurement aid (ncu target): the SURVEY 8f kernels once each on K1-sized inputs -- stereo association, the three windowed
matchers' candidate kernels, BoW descent and pair distances, distinctive descriptors, grey conversion, rotation filter."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multimot_track_b200 as orb
from bench import next_rows
from multimot_track_b200.synth import bow_match_case, local_points_case, value_noise_frame

params = (2000, 1.2, 8, 20, 7)
print({k: v for k, v in next_rows(orb, np, params, 0, 375, 1242).items() if k != "note"})
ext = orb.ORBextractor(*params)
img = value_noise_frame(0, 375, 1242)
k, d = ext(img)
_, d2 = ext(value_noise_frame(1, 375, 1242))
m = orb.ORBmatcher(0.8, True, extractor=ext)
m.SearchLocalPoints(local_points_case(1, k, d, ext.GetScaleFactors(), 3.0, 0.8, d2[:700]))
m.SearchByBoW(bow_match_case(1, k, d, 0.7, d2[:600]))
ext.extract_color(np.dstack([img, img, img]), rgb=True)
idx, d1, dd2, acc, hist, top3 = m.match_oriented(d, k["angle"], d2, np.zeros(len(d2), np.float32), 100, 0.9)
print("done", int(acc.sum()))
