"""Measurement aid: the 5000 x 5000 device-resident matcher call of bench.py, alone (ms per call, POPC.32 equivalents / s)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multimot_track_b200 as orb
ext = orb.ORBextractor(1000, 1.2, 1, 20, 7)
dA = torch.from_numpy(np.random.default_rng(1000).integers(0, 256, (5000, 32), dtype=np.uint8)).cuda()
dB = torch.from_numpy(np.random.default_rng(1001).integers(0, 256, (5000, 32), dtype=np.uint8)).cuda()
o = [torch.zeros(5000, dtype=torch.int32, device="cuda") for _ in range(3)]
lib = ext._lib
call = lambda: lib.orbx_match_device(ext._h, ctypes.c_void_p(dA.data_ptr()), 5000, ctypes.c_void_p(dB.data_ptr()), 5000, 50, 0.9,
                                     ctypes.c_void_p(o[0].data_ptr()), ctypes.c_void_p(o[1].data_ptr()), ctypes.c_void_p(o[2].data_ptr()), None)
st = torch.cuda.ExternalStream(ext.stream)
for _ in range(20):
    call()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(st):
    e0.record()
    for _ in range(200):
        call()
    e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 200
print("matcher 5000x5000: %.4f ms per call, %.1f M queries/s, %.2f T POPC.32-equivalents/s" % (ms, 5000 / ms / 1e3, 5000 * 5000 * 8 / ms / 1e9))
