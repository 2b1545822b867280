"""Measurement aid: the 5000 x 5000 matcher of BASELINE.json configs[3] on extracted descriptors, a few calls (ncu target)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multimot_track_b200 as orb
from multimot_track_b200.synth import value_noise_frame

ext = orb.ORBextractor(5000, 1.2, 8, 20, 7)
d = [np.ascontiguousarray(np.concatenate([ext(value_noise_frame(s, 1080, 1920))[1]] * 2)[:5000]) for s in (0, 1)]
dA, dB = torch.from_numpy(d[0]).cuda(), torch.from_numpy(d[1]).cuda()
o = [torch.zeros(5000, dtype=torch.int32, device="cuda") for _ in range(3)]
acc = torch.zeros(5000, dtype=torch.uint8, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(n):
    ext._lib.orbx_match_device(ext._h, ctypes.c_void_p(dA.data_ptr()), 5000, ctypes.c_void_p(dB.data_ptr()), 5000, 50, 0.9,
                               ctypes.c_void_p(o[0].data_ptr()), ctypes.c_void_p(o[1].data_ptr()), ctypes.c_void_p(o[2].data_ptr()), ctypes.c_void_p(acc.data_ptr()))
ext.sync()
print("accepted", int(acc.sum()))
