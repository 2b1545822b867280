#!/usr/bin/env python3
"""bench.py -- ORB extract+describe throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload k1|k2|k4]

A "step" is one pass of the hot path (pyramid -> FAST -> octree -> IC_Angle -> blur ->
rBRIEF, src/ORBextractor.cc:1046-1109 of the reference) over one batch of synthetic
frames.  Default workload = BASELINE.json configs[1]: 1242x375 u8, 8 levels x1.2, 2000
features, batch 32 per GPU.  One process per GPU (torchrun for N > 1); frames are sharded
frame-parallel, the data path has no collective, so scaling is "weak" (fixed work per GPU).

  value     frames/s of the whole job, inputs resident in HBM when the clock starts,
            results (keypoints + descriptors) copied back to pinned host memory inside
            the timed region; six handles (streams) keep consecutive batches in flight
  e2e       the same through the host-buffer entry point: H2D of every frame from pinned
            host memory + D2H of the results inside the timed region
  roofline  the dominant kernel: algorithmic bytes per launch / its mean
            device time, measured with CUDA events on the kernel's own stream
  cpu_baseline  the reference's CPU extractor (oracle/_ref, compiled from the reference's own
            ORBextractor.cc) on all host cores, bounded sample, rank 0 at N=1 only

`--impl reference` times only that CPU implementation (the reference arm).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (height, width, nfeatures, nlevels, batch, description)  -- BASELINE.json configs[1], [2], [4]
    "k1": (375, 1242, 2000, 8, 32, "synthetic 1242x375 u8, 8 levels x1.2, 2000 features, batch 32 per GPU"),
    "k2": (1080, 1920, 5000, 8, 64, "synthetic 1920x1080 u8, 8 levels x1.2, 5000 features, batch 64 per GPU"),
    "k4": (2160, 3840, 10000, 12, 8, "synthetic 3840x2160 u8, 12 levels x1.2, 10000 features, batch 8 per GPU"),
}
POOL_DISTINCT = 64
METRIC = "orb_extract_describe_frames_per_s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_pool(h, w):
    from multimot_track_b200.synth import frame_pool
    return frame_pool(h, w, POOL_DISTINCT, 0)


def algorithmic_bytes(level_sizes, C, K):
    """per-frame algorithmic bytes by stage (P = sum of level pixels, C candidates, K keypoints)."""
    px = [w * h for w, h in level_sizes]
    P, P0, Plast = sum(px), px[0], px[-1]
    st = {"pyramid": P0 + (P - Plast) + P, "fast": P + 8 * C, "octree": 8 * C + 12 * K,
          "orient_desc": min(749 * K, P) + 4 * K + min(512 * K, P) + 80 * K, "blur": 2 * P}
    st["total"] = sum(st.values())
    return st


def run_reference(args, H, W, nfeat, nlev, batch, desc):
    """Reference arm: the reference's own CPU ORBextractor (oracle/_ref) with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle.oracle import Oracle, RefExtractor
    cores = os.cpu_count() or 1
    params = (nfeat, 1.2, nlev, 20, 7)
    pool = make_pool(H, W)
    per_step = max(1, min(POOL_DISTINCT, cores))          # bounded sample: at most one frame per core per step
    frames = np.ascontiguousarray(pool[:per_step])
    if RefExtractor.available("asis"):
        kind, fn = "reference", lambda: RefExtractor.extract_many(params, frames, cores, "asis")
    else:
        kind, fn = "port", lambda: Oracle.extract_many(params, frames, cores)
    for _ in range(args.warmup):
        fn()
    t = 0.0
    for _ in range(args.steps):
        s, _ = fn()
        t += s
    fps = per_step * args.steps / t
    sample = "%d frames/step x %d steps of the %s pool, one extractor per thread" % (per_step, args.steps, desc.split(",")[0])
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": desc, "frames_per_step": per_step, "note": "CPU reference; GPUs unused"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_numa_node(torch, local):
    """Multi-rank runs: pin this process (and the pinned host buffers it allocates afterwards) to the CPUs of the NUMA
    node its GPU hangs off, so launches and the D2H of the results do not cross sockets.  Best effort."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="k1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--handles", type=int, default=6)
    args = ap.parse_args()
    H, W, nfeat, nlev, batch, desc = WORKLOADS[args.workload]
    if os.environ.get("ORBX_BENCH_BATCH"):                            # experiment knob: another batch size for the named shape
        batch = int(os.environ["ORBX_BENCH_BATCH"])
        desc = desc.rsplit(",", 1)[0] + ", batch %d per GPU" % batch

    if args.impl == "reference":
        return run_reference(args, H, W, nfeat, nlev, batch, desc)

    import numpy as np
    import torch
    import torch.distributed as dist
    import multimot_track_b200 as orb

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    # the pinned host buffers (first touch) and the submitting thread go to the CPUs next to the GPU; the CPU baseline later
    # gets the process's original affinity back
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_cpus = bind_to_gpu_numa_node(torch, local) if not os.environ.get("ORBX_NO_NUMA_BIND") else None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")       # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.warmup < 3:
        args.warmup = 3
    orb.load_library()

    # ---- inputs: 64 distinct frames, replicated in HBM until the pool exceeds L2 (126 MB)
    pool = make_pool(H, W)
    pitch = (W + 63) // 64 * 64
    reps = max(1, -(-(160 << 20) // (POOL_DISTINCT * H * pitch)))
    reps = max(reps, -(-2 * batch // POOL_DISTINCT))
    nslots = POOL_DISTINCT * reps
    dpool = torch.zeros((nslots, H, pitch), dtype=torch.uint8, device="cuda")
    src = torch.from_numpy(pool).cuda()
    for r in range(reps):
        dpool[r * POOL_DISTINCT:(r + 1) * POOL_DISTINCT, :, :W] = src
    del src
    hpool = torch.from_numpy(pool).pin_memory()                       # e2e inputs: pinned host memory
    hnp = hpool.numpy()
    torch.cuda.synchronize()
    nbatches = nslots // batch

    params = (nfeat, 1.2, nlev, 20, 7)
    handles = [orb.ORBextractor(*params, device_id=local, max_width=W, max_height=H, max_batch=batch) for _ in range(max(1, args.handles))]
    cap = handles[0].max_keypoints(W, H)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def submit_dev(hd, step):
        # rank r owns a contiguous block of every global batch (frame-sharded, no exchange between ranks)
        b = (step * world + rank) % nbatches
        hd.submit_device(dpool[b * batch].data_ptr(), batch, W, H, pitch, H * pitch)

    def submit_host(hd, step):
        b = ((step * world + rank) * batch) % POOL_DISTINCT
        idx = [(b + j) % POOL_DISTINCT for j in range(batch)]
        hd.submit_host([hnp[i] for i in idx])

    def run(submit, steps, first_step=0):
        """Pipelined loop: step s goes to handle s % n; a handle is collected right before it is reused."""
        nh = len(handles)
        pending = [False] * nh
        kp_total = 0
        for s in range(first_step, first_step + steps):
            i = s % nh
            if pending[i]:
                _, _, n = handles[i].collect_view()
                kp_total += int(n.sum())
            submit(handles[i], s)
            pending[i] = True
        for i in range(nh):
            if pending[i]:
                _, _, n = handles[i].collect_view()
                kp_total += int(n.sum())
        return kp_total

    def timed(submit, steps, warmup):
        run(submit, warmup)
        barrier()
        l0 = sum(h.launch_count for h in handles)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record()
        kp = run(submit, steps, first_step=warmup)
        torch.cuda.synchronize()
        ev1.record()
        barrier()
        t1 = time.time()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        nl = sum(h.launch_count for h in handles) - l0
        if world > 1:                                  # our kernels launched by all ranks inside the timed region
            c = torch.tensor([nl], dtype=torch.int64, device="cuda")
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            nl = int(c.item())
        return ms, kp, nl, (t0, t1)

    # ---- value: device-resident inputs
    # clocks / throttle reasons: rank 0 samples its own GPU (one nvidia-smi poller per job is enough)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    ms_dev, kp_dev, launches, (t0, t1) = timed(submit_dev, args.steps, args.warmup)
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- e2e: host buffers through the plugin entry point
    ms_e2e, kp_e2e, _, _ = timed(submit_host, args.steps, args.warmup)

    frames_job = batch * args.steps * world
    value = frames_job / (ms_dev * 1e-3)
    e2e = frames_job / (ms_e2e * 1e-3)

    # ---- per-stage device time (one handle, profiling events on its own stream)
    hd = handles[0]
    hd.set_profiling(True)
    acc, reps_prof = {}, 12
    for s in range(reps_prof + 2):
        submit_dev(hd, s)
        _, _, n = hd.collect_view()
        if s >= 2:
            for k, v in hd.stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v / reps_prof
    hd.set_profiling(False)
    kps_v, desc_v, n_v = None, None, n
    K = float(n.mean())
    C = float(np.mean([sum(len(hd.candidates(l, f)) for l in range(nlev)) for f in range(min(4, batch))]))
    level_sizes = [hd.level_size(l) for l in range(nlev)]
    alg = algorithmic_bytes(level_sizes, C, K)
    peak, peak_src = peaks()
    kernel_stages = {k: v for k, v in acc.items() if k in alg}
    dom = max(kernel_stages, key=kernel_stages.get)
    dom_bytes = alg[dom] * batch
    achieved = dom_bytes / (kernel_stages[dom] * 1e-3) / 1e9
    step_ms_single = sum(acc.values())
    kname = {"pyramid": "k_resize (x%d levels)" % (nlev - 1), "fast": "k_fast_fused",
             "octree": "k_octree", "orient_desc": "k_orient_desc", "blur": "k_blur"}[dom]
    traffic = None                      # dram__bytes_read+write per launch from the committed `ncu --set full` capture
    tf = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if os.path.exists(tf):
        traffic = json.load(open(tf)).get(args.workload, {}).get(kname)
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "whole_step": {"algorithmic_bytes_per_frame": alg["total"], "frac_of_peak_at_value": alg["total"] * value / world / 1e9 / peak,
                               "frac_of_8000GBs_nominal": alg["total"] * value / world / 1e9 / 8000.0},
                "stage_ms": {k: round(v, 4) for k, v in acc.items()},
                "stage_frac_of_peak": {k: alg[k] * batch / (v * 1e-3) / 1e9 / peak for k, v in kernel_stages.items() if v > 0},
                "candidates_per_frame": C, "keypoints_per_frame": K}

    # ---- matcher (BASELINE.json configs[3]): 5000 x 5000 x 256 bit, device-resident, TH_LOW + 0.9 ratio
    matcher = None
    try:
        rng = np.random.default_rng(1000)
        dA = torch.from_numpy(rng.integers(0, 256, (5000, 32), dtype=np.uint8)).cuda()
        dB = torch.from_numpy(np.random.default_rng(1001).integers(0, 256, (5000, 32), dtype=np.uint8)).cuda()
        o_idx = torch.zeros(5000, dtype=torch.int32, device="cuda"); o_d1 = torch.zeros_like(o_idx); o_d2 = torch.zeros_like(o_idx)
        lib = hd._lib
        st = torch.cuda.ExternalStream(hd.stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import ctypes
        call = lambda: lib.orbx_match_device(hd._h, ctypes.c_void_p(dA.data_ptr()), 5000, ctypes.c_void_p(dB.data_ptr()), 5000, 50, 0.9,
                                             ctypes.c_void_p(o_idx.data_ptr()), ctypes.c_void_p(o_d1.data_ptr()), ctypes.c_void_p(o_d2.data_ptr()), None)
        for _ in range(5):
            call()
        hd.sync()
        e0.record(st)
        for _ in range(50):
            call()
        e1.record(st)
        hd.sync()
        mms = e0.elapsed_time(e1) / 50
        matcher = {"workload": "5000x5000 256-bit brute force, TH_LOW=50, ratio 0.9, device-resident", "ms_per_call": mms,
                   "queries_per_s": 5000 / (mms * 1e-3), "pair_distances_per_s": 25e6 / (mms * 1e-3),
                   "popc32_per_s": 2e8 / (mms * 1e-3)}
    except Exception as ex:                                            # the headline metric must still print
        matcher = {"error": repr(ex)}

    # ---- the "next" rows of SURVEY 8f, timed end to end through the C ABI (host wall clock, results on the host)
    extras = None
    if rank == 0:
        try:
            import tempfile
            from multimot_track_b200.synth import stereo_pair, write_synthetic_vocabulary
            eL, eR = orb.ORBextractor(*params, device_id=local), orb.ORBextractor(*params, device_id=local)
            Ls, Rs = stereo_pair(3, H, W)
            kL, dL = eL(Ls); kR, dR = eR(Rs)
            t0 = time.perf_counter()
            for _ in range(50):
                eR(Rs)
            t_single = (time.perf_counter() - t0) / 50
            eL.stereo_match(eR, 386.1448)
            t0 = time.perf_counter()
            for _ in range(20):
                ur, dp, di, kept = eL.stereo_match(eR, 386.1448)
            t_stereo = (time.perf_counter() - t0) / 20
            with tempfile.TemporaryDirectory() as td:
                vp = os.path.join(td, "voc.txt")
                write_synthetic_vocabulary(vp, 10, 4, seed=1, seeds=dL)
                voc = orb.ORBVocabulary(eL)
                voc.loadFromTextFile(vp)
            voc.transform(dL, 4)
            t0 = time.perf_counter()
            for _ in range(20):
                voc.transform(dL, 4)
            t_bow = (time.perf_counter() - t0) / 20
            rng2 = np.random.default_rng(5)
            sizes = rng2.integers(2, 40, 5000)
            off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
            rows = dL[rng2.integers(0, len(dL), off[-1])]
            eL.distinctive_descriptors(rows, off)
            t0 = time.perf_counter()
            for _ in range(10):
                eL.distinctive_descriptors(rows, off)
            t_dd = (time.perf_counter() - t0) / 10
            from multimot_track_b200.synth import projection_case
            mt = orb.ORBmatcher(0.9, True, extractor=eL)
            pc = projection_case(1, kL, dL, eL.GetScaleFactors(), 15.0, False, 0.0, dR[:700])
            mt.SearchByProjection(pc)
            t0 = time.perf_counter()
            for _ in range(20):
                _, nproj = mt.SearchByProjection(pc)
            t_proj = (time.perf_counter() - t0) / 20
            from multimot_track_b200.synth import initialization_case
            ic = initialization_case(1, kL, dL, 100, 0.9, (6.0, -3.0), dR[:600])
            mt.SearchForInitialization(ic)
            t0 = time.perf_counter()
            for _ in range(20):
                _, _, ninit = mt.SearchForInitialization(ic)
            t_init = (time.perf_counter() - t0) / 20
            from multimot_track_b200.synth import local_points_case
            lc = local_points_case(1, kL, dL, eL.GetScaleFactors(), 3.0, 0.8, dR[:700])
            mt.SearchLocalPoints(lc)
            t0 = time.perf_counter()
            for _ in range(20):
                _, nloc = mt.SearchLocalPoints(lc)
            t_loc = (time.perf_counter() - t0) / 20
            from multimot_track_b200.synth import bow_match_case
            bc = bow_match_case(1, kL, dL, 0.7, dR[:600])
            mt.SearchByBoW(bc)
            t0 = time.perf_counter()
            for _ in range(20):
                _, nbow = mt.SearchByBoW(bc)
            t_bowm = (time.perf_counter() - t0) / 20
            extras = {"search_by_bow": {"workload": "ORBmatcher::SearchByBoW(pKF, F), %d x %d features in %d / %d nodes" % (len(kL), len(bc["f_desc"]), len(bc["kf_nodes"]), len(bc["f_nodes"])),
                                        "ms_per_call": 1e3 * t_bowm, "nmatches": int(nbow)},
                      "search_local_points": {"workload": "ORBmatcher::SearchByProjection(F, vpMapPoints, th=3), %d map points x %d features" % (len(lc["proj"]), len(lc["xy"])),
                                              "ms_per_call": 1e3 * t_loc, "nmatches": int(nloc)},
                      "search_for_initialization": {"workload": "ORBmatcher::SearchForInitialization, %d x %d keypoints, window 100, ratio 0.9" % (len(kL), len(ic["xy2"])),
                                                    "ms_per_call": 1e3 * t_init, "nmatches": int(ninit)},
                      "search_by_projection": {"workload": "ORBmatcher::SearchByProjection(CurrentFrame, LastFrame), %d map points x %d features, th 15" % (len(kL), len(pc["cur_xy"])),
                                               "ms_per_call": 1e3 * t_proj, "nmatches": int(nproj)},
                      "single_frame_call": {"workload": "ORBextractor::operator() on one %dx%d host frame (H2D, extract, D2H), the shape Frame::ExtractORB calls" % (W, H),
                                            "ms_per_frame": 1e3 * t_single},
                      "stereo_match": {"workload": "Frame::ComputeStereoMatches, %d x %d keypoints, %dx%d pair" % (len(kL), len(kR), W, H),
                                       "ms_per_pair": 1e3 * t_stereo, "matches_kept": int(kept)},
                      "bow_transform": {"workload": "ORBVocabulary::transform, %d descriptors, synthetic k=10 L=4 tree" % len(dL), "ms_per_frame": 1e3 * t_bow},
                      "distinctive_descriptors": {"workload": "5000 map points, %d observations" % int(off[-1]), "ms_per_batch": 1e3 * t_dd},
                      "note": "host wall clock through the synchronous C-ABI calls, H2D / D2H of the small arrays included"}
        except Exception as ex:
            extras = {"error": repr(ex)}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's ORBextractor on all host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.oracle import Oracle, RefExtractor
        if affinity0 is not None:
            os.sched_setaffinity(0, affinity0)                          # all host cores for the reference
        cores = os.cpu_count() or 1
        chunk = np.ascontiguousarray(pool[:max(1, min(POOL_DISTINCT, cores))])
        if RefExtractor.available("asis"):
            kind, fn = "reference", lambda: RefExtractor.extract_many(params, chunk, cores, "asis")
        else:
            kind, fn = "port", lambda: Oracle.extract_many(params, chunk, cores)
        fn()
        tot_s, tot_f = 0.0, 0
        while tot_s < 12.0 and tot_f < 200000:
            s, _ = fn()
            tot_s += s; tot_f += len(chunk)
        cpu = {"value": tot_f / tot_s, "unit": "frames/s", "cores": cores, "kind": kind,
               "sample": "%d frames of the same pool in chunks of %d, %.1f s wall, one extractor instance per thread" % (tot_f, len(chunk), tot_s),
               "ms_per_frame_per_core": 1e3 * tot_s * min(cores, len(chunk)) / tot_f}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": desc, "nfeatures": nfeat, "nlevels": nlev, "scale_factor": 1.2, "ini_th_fast": 20, "min_th_fast": 7,
                           "batch_per_gpu": batch, "global_batch": batch * world, "sharding": "frame-parallel, no collective", "numa_bind": numa_cpus,
                           "handles_in_flight": len(handles),
                           "l2": "inputs larger than L2: %d frame slots = %.0f MB in HBM, walked cyclically" % (nslots, nslots * H * pitch / 2 ** 20),
                           "input_row_pitch": pitch, "keypoints_per_step": kp_dev / max(1, args.steps)},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": batch * H * W, "d2h_bytes_per_step": batch * (cap * 60 + 4),
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "matcher": matcher, "next_rows": extras}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
