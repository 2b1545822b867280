"""Measurement aid: per-phase clock cycles of k_octree for frame 0 of a batch (library built with -DORBX_OCT_TIMING,
passed through ORBX_LIBRARY).  usage: ORBX_LIBRARY=.../liborbx_timing.so python scripts/exp_octree_timing.py [k1|k4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import multimot_track_b200 as orb
from multimot_track_b200.synth import value_noise_frame
wl = sys.argv[1] if len(sys.argv) > 1 else "k1"
H, W, N, L, B = {"k1": (375, 1242, 2000, 8, 32), "k4": (2160, 3840, 10000, 12, 4)}[wl]
ext = orb.ORBextractor(N, 1.2, L, 20, 7)
frames = [value_noise_frame(s, H, W) for s in range(min(B, 4))] * (B // min(B, 4))
for rep in range(2):
    print("--- rep", rep, flush=True)
    ext.extract_batch(frames) if hasattr(ext, "extract_batch") else [ext(f) for f in frames]
