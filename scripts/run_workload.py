"""Measurement aid (ncu target): a few device-resident batches of one BASELINE.json workload through one handle, nothing else.
usage: python scripts/run_workload.py k1|k2|k4 [nbatches]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimot_track_b200 as orb
from bench import WORKLOADS
from multimot_track_b200.synth import frame_pool

wl = sys.argv[1] if len(sys.argv) > 1 else "k1"
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H, W, nfeat, nlev, batch, desc = WORKLOADS[wl]
pool = frame_pool(H, W, min(batch, 16), 0)
pitch = (W + 63) // 64 * 64
dev = torch.zeros((batch, H, pitch), dtype=torch.uint8, device="cuda")
for f in range(batch):
    dev[f, :, :W] = torch.from_numpy(pool[f % len(pool)]).cuda()
torch.cuda.synchronize()
ext = orb.ORBextractor(nfeat, 1.2, nlev, 20, 7, max_width=W, max_height=H, max_batch=batch)
for _ in range(nb):
    ext.submit_device(dev.data_ptr(), batch, W, H, pitch, H * pitch)
    _, _, n = ext.collect_view()
print(desc, "keypoints per frame", float(n.mean()))
