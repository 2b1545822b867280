"""Measurement aid: throughput of the device-resident K1 loop against the number of handles in flight, the host-side cost of a
submit, and the marginal cost of each stage in the overlapped pipeline (ORBX_DEBUG_SKIP).  Not a bench: prints a table."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import multimot_track_b200 as orb          # run with ORBX_LIBRARY=multimot_track_b200/liborbx_dbg.so (make -C multimot_track_b200/csrc dbg): the release library has no ORBX_DEBUG_SKIP
from bench import make_pool, POOL_DISTINCT

from bench import WORKLOADS
WL = os.environ.get("EXP_WORKLOAD", "k1")
H, W, NFEAT, NLEV, batch, _ = WORKLOADS[WL]
pool = make_pool(H, W)
pitch = (W + 63) // 64 * 64
reps = 6 if WL == "k1" else 1
nslots = POOL_DISTINCT * reps
dpool = torch.zeros((nslots, H, pitch), dtype=torch.uint8, device="cuda")
src = torch.from_numpy(pool).cuda()
for r in range(reps):
    dpool[r * POOL_DISTINCT:(r + 1) * POOL_DISTINCT, :, :W] = src
nb = nslots // batch
NH = int(os.environ.get("EXP_MAX_HANDLES", "12"))
handles = [orb.ORBextractor(NFEAT, 1.2, NLEV, 20, 7, device_id=0, max_width=W, max_height=H, max_batch=batch) for _ in range(NH)]


def run(nh, steps, first=0):
    pend = [False] * nh
    for s in range(first, first + steps):
        i = s % nh
        if pend[i]:
            handles[i].collect_view()
        handles[i].submit_device(dpool[(s % nb) * batch].data_ptr(), batch, W, H, pitch, H * pitch)
        pend[i] = True
    for i in range(nh):
        if pend[i]:
            handles[i].collect_view()


def timed(nh, steps=200 if WL == "k1" else 60):
    run(nh, 20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(nh, steps, 20)
    torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


for nh in [int(x) for x in os.environ.get("EXP_HANDLES", "1,2,3,4,6,8,12").split(",")]:
    print("handles %2d: %.4f ms/step" % (nh, timed(nh)), flush=True)
# host cost of a submit: enqueue on idle handles without waiting
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(NH):
    handles[i].submit_device(dpool[(i % nb) * batch].data_ptr(), batch, W, H, pitch, H * pitch)
t1 = time.perf_counter()
for i in range(NH):
    handles[i].collect_view()
print("host submit: %.1f us per batch" % ((t1 - t0) / NH * 1e6), flush=True)
base = timed(6)
print("full      : %.4f ms/step" % base)
for name, bit in (("pyramid", 1), ("blur", 2), ("fast", 4), ("octree", 8), ("orient", 16), ("d2h", 32), ("all-but-fast", 59), ("all-but-octree", 55)):
    os.environ["ORBX_DEBUG_SKIP"] = str(bit)
    t = timed(6)
    print("skip %-14s: %.4f ms/step (marginal %.4f)" % (name, t, base - t), flush=True)
os.environ["ORBX_DEBUG_SKIP"] = "0"
