"""Measurement aid (GPU, debug-knob build): the floor of any "pre-test every pixel, exact score only for survivors" FAST kernel.
ORBX_FAST_BOUND=1 makes k_fast_fused compute the cheap upper bound of the score (8 opposing ring pairs, 26 packed min / max)
instead of the exact score (72) for EVERY pixel, with the same staging, window loads, NMS and list machinery -- the work a
pre-test scheme cannot avoid.  Results are meaningless in that mode; only the time matters.
usage: ORBX_LIBRARY=multimot_track_b200/liborbx_dbg.so python scripts/exp_fast_floor.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimot_track_b200 as orb
from bench import make_pool, POOL_DISTINCT

H, W, batch = 375, 1242, 32
pool = make_pool(H, W)
pitch = (W + 63) // 64 * 64
dpool = torch.zeros((POOL_DISTINCT * 6, H, pitch), dtype=torch.uint8, device="cuda")
src = torch.from_numpy(pool).cuda()
for r in range(6):
    dpool[r * POOL_DISTINCT:(r + 1) * POOL_DISTINCT, :, :W] = src
nb = dpool.shape[0] // batch
handles = [orb.ORBextractor(2000, 1.2, 8, 20, 7, device_id=0, max_width=W, max_height=H, max_batch=batch) for _ in range(6)]


def loop(steps, first=0):
    pend = [False] * 6
    for s in range(first, first + steps):
        i = s % 6
        if pend[i]:
            handles[i].collect_view()
        handles[i].submit_device(dpool[(s % nb) * batch].data_ptr(), batch, W, H, pitch, H * pitch)
        pend[i] = True
    for i in range(6):
        if pend[i]:
            handles[i].collect_view()


def timed(steps=300):
    loop(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop(steps, 20); torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def stage(name):
    h = handles[0]
    h.set_profiling(True)
    acc = 0.0
    for s in range(14):
        h.submit_device(dpool[(s % nb) * batch].data_ptr(), batch, W, H, pitch, H * pitch)
        h.collect_view()
        if s >= 2:
            acc += h.stage_ms()[name] / 12
    h.set_profiling(False)
    return acc


for mode, label in ((0, "exact score (72 packed min/max per pixel pair)"), (1, "upper bound only (26 packed min/max per pixel pair)")):
    os.environ["ORBX_FAST_BOUND"] = str(mode)
    print("%-52s  FAST stage alone %.4f ms   6-handle step %.4f ms" % (label, stage("fast"), timed()), flush=True)
os.environ["ORBX_FAST_BOUND"] = "0"
os.environ["ORBX_DEBUG_SKIP"] = "4"
print("%-52s  FAST stage alone %.4f ms   6-handle step %.4f ms" % ("FAST left out (ORBX_DEBUG_SKIP=4)", 0.0, timed()), flush=True)
