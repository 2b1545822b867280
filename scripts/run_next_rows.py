"""Measurement aid (ncu target): the SURVEY 8f kernels ONCE each on K1-sized inputs -- stereo association, the three windowed
matchers' candidate kernels, BoW descent and pair distances, distinctive descriptors, grey conversion, rotation filter, border."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multimot_track_b200 as orb
from multimot_track_b200.synth import (bow_match_case, initialization_case, local_points_case, projection_case, stereo_pair,
                                       value_noise_frame, write_synthetic_vocabulary)

params = (2000, 1.2, 8, 20, 7)
eL, eR = orb.ORBextractor(*params), orb.ORBextractor(*params)
L, R = stereo_pair(3, 375, 1242)
kL, dL = eL(L); kR, dR = eR(R)
eL.stereo_match(eR, 386.1448)                                              # k_stereo_match, k_stereo_median
m = orb.ORBmatcher(0.9, True, extractor=eL)
sc = eL.GetScaleFactors()
m.SearchByProjection(projection_case(1, kL, dL, sc, 15.0, False, 0.0, dR[:700]))          # k_project_candidates
m.SearchForInitialization(initialization_case(1, kL, dL, 100, 0.9, (6.0, -3.0), dR[:600]))   # k_window_candidates
m.SearchLocalPoints(local_points_case(1, kL, dL, sc, 3.0, 0.8, dR[:700]))                  # k_local_candidates
m.SearchByBoW(bow_match_case(1, kL, dL, 0.7, dR[:600]))                                    # k_bow_pair_distances
with tempfile.TemporaryDirectory() as td:
    vp = os.path.join(td, "voc.txt")
    write_synthetic_vocabulary(vp, 10, 4, seed=1, seeds=dL)
    voc = orb.ORBVocabulary(eL)
    voc.loadFromTextFile(vp)
voc.transform(dL, 4)                                                       # k_bow_descent
rng = np.random.default_rng(5)
sizes = rng.integers(2, 40, 5000)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
eL.distinctive_descriptors(dL[rng.integers(0, len(dL), off[-1])], off)     # k_distinctive
eL.extract_color(np.dstack([L, L, L]), rgb=True)                           # k_gray
m.match_oriented(dL, kL["angle"], dR, kR["angle"], 100, 0.9)               # k_match_partial, k_rotation_filter
eL.pyramid_level(0, with_border=True)                                      # k_pad_reflect101
print("done")
