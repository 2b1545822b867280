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
            the timed region; the native dispatcher (orbx_pool_*, one worker thread per GPU)
            keeps six batches in flight on six handles / streams
  e2e       the same through the host-buffer entry point: H2D of every frame from pinned
            host memory + D2H of the results inside the timed region
  roofline  the dominant kernel: algorithmic bytes per launch / its mean
            device time, measured with CUDA events on the kernel's own stream
  cpu_baseline  two CPU legs on all host cores, bounded samples, rank 0 at N=1 only: the reference's own
            ORBextractor.cc compiled against the oracle's OpenCV-primitive shim (oracle/_ref, T threads), and the
            Python + cv2 4.13 restatement as T single-threaded processes; `kind` names the faster one
  parity    after the timed region one more batch goes through both entry points and every frame is compared with
            the CPU oracle (parity_checked)

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
LANES = {"k1": 1, "k2": 1, "k4": 1}        # dispatcher workers per GPU (sub-batches per batch), measured per workload
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

    def __init__(self, gpu_index, cpus=None):
        self.rows, self.proc, self.gpu, self.cpus = [], None, gpu_index, cpus

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                                         preexec_fn=(lambda: os.sched_setaffinity(0, self.cpus)) if self.cpus else None)   # not on the bench's own CPUs
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
            "config": {"workload": desc, "nfeatures": nfeat, "nlevels": nlev, "scale_factor": 1.2, "ini_th_fast": 20, "min_th_fast": 7,
                       "batch_per_gpu": batch, "frames_per_step": per_step,
                       "note": "CPU reference (the reference's own ORBextractor.cc, %s); GPUs unused; a step is a bounded sample of the batch" % kind},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def bind_to_gpu_cpus(torch, local, world_local):
    """Multi-rank runs: pin this process (submitting thread, native worker thread, and the pinned host buffers it first-touches
    afterwards) to CPUs next to its GPU.  /sys gives the GPU's local CPU list; when several ranks' GPUs share one list (a single
    NUMA node, or a VM that exposes one) the list is split evenly between the local ranks so they do not migrate onto each other.
    Best effort; returns the CPU set as text."""
    try:
        p = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        ids = []
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(ids) & os.sched_getaffinity(0)) or sorted(os.sched_getaffinity(0))
        if world_local > 1 and len(allowed) >= 2 * world_local:
            per = len(allowed) // world_local
            allowed = allowed[local * per:(local + 1) * per]
        os.sched_setaffinity(0, set(allowed))
        return "%d-%d (%d cpus)" % (allowed[0], allowed[-1], len(allowed))
    except Exception:
        return None


def popc_peak():
    """Measured POPC.32 rate of this GPU (scripts/ubench/popc, built by __graft_entry__.build()): the matcher's roofline."""
    exe = os.path.join(ROOT, "scripts", "ubench", "popc")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="k1", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the matcher / next-row timings (profiling runs)")
    ap.add_argument("--cpu-bind", default="none", choices=["auto", "none"], help="multi-rank runs: pin each rank to its own slice of the GPU's CPUs")
    ap.add_argument("--handles", type=int, default=6, help="batches in flight per GPU (depth of the native dispatcher)")
    ap.add_argument("--batch", type=int, default=0, help="experiment: another batch size for the named shape")
    ap.add_argument("--fast-tma", action="store_true", help="experiment: the persistent TMA-staged variant of the FAST kernel (ORBX_OPT_FAST_TMA); named in config")
    ap.add_argument("--full-records", action="store_true", help="device-resident loop with 28-byte cv::KeyPoint records instead of the 12-byte compact ones")
    ap.add_argument("--lanes", type=int, default=0, help="dispatcher workers per GPU: a batch is split into that many contiguous sub-batches, each with its own "
                                                          "worker thread, handles and streams (0 = the workload's default)")
    args = ap.parse_args()
    t_start = time.time()
    H, W, nfeat, nlev, batch, desc = WORKLOADS[args.workload]
    if args.batch > 0:
        batch = args.batch
        desc = desc.rsplit(",", 1)[0] + ", batch %d per GPU" % batch

    if args.impl == "reference":
        return run_reference(args, H, W, nfeat, nlev, batch, desc)

    import numpy as np
    import torch
    import torch.distributed as dist
    import multimot_track_b200 as orb
    from multimot_track_b200.sharding import job_throughput, shard_bounds

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    world_local = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    # the pinned host buffers (first touch), the submitting thread and the dispatcher's worker thread go to CPUs next to the GPU;
    # the CPU baseline later gets the process's original affinity back
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    cpu_bind = bind_to_gpu_cpus(torch, local, world_local) if (world > 1 and affinity0 is not None and args.cpu_bind == "auto") else None
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
    depth = max(1, args.handles)
    # the native frame-sharded dispatcher with this rank's GPU as its only device: one C++ worker thread issues the copies and
    # launches, `depth` handles (streams) keep consecutive batches in flight; Python only queues tickets and collects views
    lanes = args.lanes if args.lanes > 0 else LANES[args.workload]
    lanes = max(1, min(lanes, batch))
    sub = [shard_bounds(batch, g, lanes) for g in range(lanes)]      # contiguous sub-batches of this rank's block, one per worker
    xpool = orb.ExtractorPool(*params, devices=[local] * lanes, depth=depth, max_width=W, max_height=H, max_batch=max(hi - lo for lo, hi in sub))
    OPT_COMPACT = orb.ORBextractor.OPT_COMPACT_KEYPOINTS
    compact = not args.full_records
    if args.fast_tma:
        xpool.set_option(orb.ORBextractor.OPT_FAST_TMA, 1)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def frames_of_step(step):
        # rank r owns the contiguous block [r*B/G, (r+1)*B/G) of every global batch of B = batch * world frames (frame-sharded,
        # no exchange between ranks); the global batches walk the pool cyclically
        lo, hi = shard_bounds(batch * world, rank, world)
        assert hi - lo == batch
        return step * world + rank

    def submit_dev(step):
        b = frames_of_step(step) % nbatches
        return xpool.submit_device([dpool[b * batch + lo].data_ptr() for lo, hi in sub], [hi - lo for lo, hi in sub], W, H, pitch, H * pitch)

    def submit_host(step):
        b = (frames_of_step(step) * batch) % POOL_DISTINCT
        return xpool.submit_host([hnp[(b + j) % POOL_DISTINCT] for j in range(batch)])

    def prime():
        """Setup, not measurement: every handle of the dispatcher sees one batch through each entry point, so lazy first-use work
        (kernel module loads, tensor-map encodes, pinned staging) is not charged to whichever step first lands on a cold handle."""
        for submit in (submit_dev, submit_host):
            for t in [submit(s) for s in range(depth)]:
                xpool.collect(t)

    def run(submit, steps, first_step=0):
        """Pipelined loop: at most `depth` tickets outstanding; the oldest is collected before the next submit."""
        tickets, kp_total = [], 0
        for s in range(first_step, first_step + steps):
            if len(tickets) == depth:
                kp_total += sum(int(sh[3].sum()) for sh in xpool.collect(tickets.pop(0)))
            tickets.append(submit(s))
        for t in tickets:
            kp_total += sum(int(sh[3].sum()) for sh in xpool.collect(t))
        return kp_total

    def timed(submit, steps, warmup):
        run(submit, warmup)
        barrier()
        l0 = xpool.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        ev0.record()
        kp = run(submit, steps, first_step=warmup)
        torch.cuda.synchronize()
        ev1.record()
        barrier()
        t1 = time.time()
        ms = ev0.elapsed_time(ev1)
        nl = xpool.launch_count() - l0
        per_rank = [ms]
        if world > 1:
            g = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
            dist.all_gather(g, torch.tensor([ms], dtype=torch.float64, device="cuda"))
            per_rank = [float(x.item()) for x in g]
            ms = max(per_rank)                                        # the job is as slow as its slowest rank
            c = torch.tensor([nl], dtype=torch.int64, device="cuda")  # our kernels launched by all ranks inside the timed region
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            nl = int(c.item())
        return ms, kp, nl, (t0, t1, per_rank)

    # ---- value: device-resident inputs
    # clocks / throttle reasons: rank 0 samples its own GPU (one nvidia-smi poller per job is enough)
    sampler = ClockSampler(local, (affinity0 - os.sched_getaffinity(0)) or None if affinity0 is not None else None) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    note = (lambda m: print("[bench %.1fs] %s" % (time.time() - t_start, m), file=sys.stderr, flush=True)) if rank == 0 else (lambda m: None)
    prime()
    xpool.set_option(OPT_COMPACT, int(compact))
    prime()
    note("pool of %d frames resident, dispatcher up and primed" % nslots)
    # device-resident loop: the results land in pinned host memory as descriptors + 12-byte compact keypoint records (lossless:
    # orbx_expand_keypoints rebuilds the cv::KeyPoint exactly; the self-check below does that for every record it compares)
    ms_dev, kp_dev, launches, (t0, t1, per_rank_dev) = timed(submit_dev, args.steps, args.warmup)
    note("device-resident loop done: %.4f ms/step" % (ms_dev / args.steps))
    clocks = sampler.stop(t0, t1) if sampler else None
    # ---- e2e: host buffers through the plugin entry point, results as full 28-byte cv::KeyPoint records + descriptors
    xpool.set_option(OPT_COMPACT, 0)
    ms_e2e, kp_e2e, _, (_, _, per_rank_e2e) = timed(submit_host, args.steps, args.warmup)
    note("host-buffer loop done: %.4f ms/step" % (ms_e2e / args.steps))

    value = job_throughput([batch * args.steps] * world, [ms_dev * 1e-3] * world)
    e2e = job_throughput([batch * args.steps] * world, [ms_e2e * 1e-3] * world)

    # ---- parity self-check: one more batch through each entry point, every frame against the CPU oracle (rank 0)
    parity = None
    if rank == 0:
        from oracle.oracle import Oracle
        import concurrent.futures as cf
        check_step = args.warmup + args.steps
        b_dev = frames_of_step(check_step) % nbatches
        idx_dev = [(b_dev * batch + j) % POOL_DISTINCT for j in range(batch)]
        b_host = (frames_of_step(check_step) * batch) % POOL_DISTINCT
        idx_host = [(b_host + j) % POOL_DISTINCT for j in range(batch)]
        need = sorted(set(idx_dev) | set(idx_host))
        if affinity0 is not None:
            os.sched_setaffinity(0, affinity0)
        workers = min(len(need), os.cpu_count() or 1)

        def oracle_chunk(chunk):
            o = Oracle(*params)
            return [(i,) + tuple(a.copy() for a in o(pool[i])) for i in chunk]
        ref = {}
        with cf.ThreadPoolExecutor(workers) as ex:
            for part in ex.map(oracle_chunk, [need[w::workers] for w in range(workers)]):
                for i, k, d in part:
                    ref[i] = (k, d)
        bad_frames = bad_bits = bits = 0
        max_angle = 0.0
        for submit, idx, cmp in ((submit_dev, idx_dev, compact), (submit_host, idx_host, False)):
            xpool.set_option(OPT_COMPACT, int(cmp))
            shards = xpool.collect(submit(check_step))
            for f, i in enumerate(idx):
                first, kps, descs, n = [sh for sh in shards if sh[0] <= f < sh[0] + len(sh[3])][0]
                k, d = kps[f - first, :n[f - first]], descs[f - first, :n[f - first]]
                if cmp:
                    k = xpool.expand_keypoints(k)
                rk, rd = ref[i]
                same = len(k) == len(rk) and all(np.array_equal(k[c], rk[c]) for c in ("x", "y", "size", "response", "octave", "class_id"))
                if same:
                    da = np.abs(k["angle"].astype(np.float64) - rk["angle"].astype(np.float64))
                    max_angle = max(max_angle, float(np.deg2rad(np.minimum(da, 360.0 - da)).max(initial=0.0)))
                    bad_bits += int(np.unpackbits(d ^ rd).sum()); bits += d.size * 8
                else:
                    bad_frames += 1
        parity = {"checked": bad_frames == 0 and max_angle <= 1e-4 and bad_bits <= 1e-4 * max(bits, 1),
                  "frames": 2 * batch, "entry_points": ["orbx_pool_submit_device", "orbx_pool_submit_host"],
                  "frames_with_keypoint_mismatch": bad_frames, "max_angle_diff_rad": max_angle,
                  "descriptor_bits_differing": bad_bits, "descriptor_bits": bits, "against": "oracle/orb_oracle.c (pinned to the reference's compiled ORBextractor.cc)"}

    note("parity self-check: %s" % (parity,))
    # ---- per-stage device time (one handle, profiling events on its own stream)
    hd = orb.ORBextractor(*params, device_id=local, max_width=W, max_height=H, max_batch=batch)
    hd.set_profiling(True)
    acc, reps_prof = {}, 12
    for s in range(reps_prof + 2):
        b = frames_of_step(s) % nbatches
        hd.submit_device(dpool[b * batch].data_ptr(), batch, W, H, pitch, H * pitch)
        _, _, n = hd.collect_view()
        if s >= 2:
            for k, v in hd.stage_ms().items():
                acc[k] = acc.get(k, 0.0) + v / reps_prof
    hd.set_profiling(False)
    K = float(n.mean())
    C = float(np.mean([sum(len(hd.candidates(l, f)) for l in range(nlev)) for f in range(min(4, batch))]))
    level_sizes = [hd.level_size(l) for l in range(nlev)]
    cap = hd.max_keypoints(W, H)
    alg = algorithmic_bytes(level_sizes, C, K)
    peak, peak_src = peaks()
    kernel_stages = {k: v for k, v in acc.items() if k in alg}
    dom = max(kernel_stages, key=kernel_stages.get)
    dom_bytes = alg[dom] * batch
    achieved = dom_bytes / (kernel_stages[dom] * 1e-3) / 1e9
    kname = {"pyramid": "k_resize_sep", "fast": "k_fast_fused", "octree": "k_octree", "orient_desc": "k_orient_desc_tma", "blur": "k_blur"}[dom]
    # dram__bytes_read + write of that kernel per launch: NOT measured in this run (a run under ncu is never a bench run); read from
    # the committed `ncu --set full` capture of this same command, when there is one for the workload
    traffic, traffic_src = None, None
    for tf in ("r2b_traffic.json", "r2_traffic.json", "r1_traffic.json"):   # newest capture that holds this workload
        tp = os.path.join(ROOT, "profiles", tf)
        if os.path.exists(tp):
            v = json.load(open(tp)).get(args.workload, {}).get(kname)
            if v is not None:
                traffic, traffic_src = v, "profiles/%s (committed ncu --set full capture of `bench.py --workload %s`, bytes per launch)" % (tf, args.workload)
                break
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                "whole_step": {"algorithmic_bytes_per_frame": alg["total"], "frac_of_peak_at_value": alg["total"] * value / world / 1e9 / peak,
                               "frac_of_8000GBs_nominal": alg["total"] * value / world / 1e9 / 8000.0,
                               "fusion_ideal_bytes_per_frame": level_sizes[0][0] * level_sizes[0][1] + sum(w * h for w, h in level_sizes) + 60 * K},
                "stage_ms": {k: round(v, 4) for k, v in acc.items()},
                "stage_frac_of_peak": {k: alg[k] * batch / (v * 1e-3) / 1e9 / peak for k, v in kernel_stages.items() if v > 0},
                "candidates_per_frame": C, "keypoints_per_frame": K}

    # ---- matcher (BASELINE.json configs[3]): 5000 x 5000 x 256 bit, device-resident; extracted descriptors (pool frames s, s+1 of the
    #      1920x1080 / 5000-feature shape, cropped to 5000 rows) at TH_LOW and TH_HIGH with the 0.9 ratio test, the random-bit kernel
    #      number, the CPU scan on all cores beside it and the measured POPC.32 peak of this GPU (integer-pipe roofline)
    matcher = None
    if rank == 0 and not args.no_extras:
        try:
            import ctypes
            from multimot_track_b200.synth import value_noise_frame
            from oracle.oracle import Oracle
            ext5k = orb.ORBextractor(5000, 1.2, 8, 20, 7, device_id=local)
            dpair = []
            for sd in (0, 1):
                d = ext5k(value_noise_frame(sd, 1080, 1920))[1]
                dpair.append(np.ascontiguousarray(np.concatenate([d] * (-(-5000 // len(d))))[:5000]))
            rnd = [np.random.default_rng(sd).integers(0, 256, (5000, 32), dtype=np.uint8) for sd in (1000, 1001)]
            o_idx = torch.zeros(5000, dtype=torch.int32, device="cuda"); o_d1 = torch.zeros_like(o_idx); o_d2 = torch.zeros_like(o_idx)
            o_acc = torch.zeros(5000, dtype=torch.uint8, device="cuda")
            lib = hd._lib
            st = torch.cuda.ExternalStream(hd.stream)

            def time_match(A, B, th):
                dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
                call = lambda: lib.orbx_match_device(hd._h, ctypes.c_void_p(dA.data_ptr()), 5000, ctypes.c_void_p(dB.data_ptr()), 5000, th, 0.9,
                                                     ctypes.c_void_p(o_idx.data_ptr()), ctypes.c_void_p(o_d1.data_ptr()), ctypes.c_void_p(o_d2.data_ptr()),
                                                     ctypes.c_void_p(o_acc.data_ptr()))
                for _ in range(5):
                    call()
                hd.sync()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(50):
                    call()
                e1.record(st)
                hd.sync()
                return e0.elapsed_time(e1) / 50, int(o_acc.sum().item()), o_idx.cpu().numpy(), o_d1.cpu().numpy()
            ms_lo, acc_lo, gi, g1 = time_match(dpair[0], dpair[1], 50)
            ms_hi, acc_hi, _, _ = time_match(dpair[0], dpair[1], 100)
            ms_rnd, acc_rnd, _, _ = time_match(rnd[0], rnd[1], 50)
            cores = os.cpu_count() or 1
            if affinity0 is not None:
                os.sched_setaffinity(0, affinity0)
            Oracle.match(dpair[0], dpair[1], 50, 0.9, threads=cores)
            tc = time.perf_counter()
            for _ in range(3):
                ci, c1, c2, ca = Oracle.match(dpair[0], dpair[1], 50, 0.9, threads=cores)
            cpu_ms = (time.perf_counter() - tc) / 3 * 1e3
            pk = popc_peak()
            mms = ms_lo
            matcher = {"workload": "5000x5000 256-bit brute force on extracted descriptors (1920x1080 pool frames 0 / 1, 5000 rows), ratio 0.9, device-resident",
                       "ms_per_call": mms, "ms_per_call_th_high": ms_hi, "ms_per_call_random_bits": ms_rnd,
                       "queries_per_s": 5000 / (mms * 1e-3), "pair_distances_per_s": 25e6 / (mms * 1e-3),
                       "accepted_th_low": acc_lo, "accepted_th_high": acc_hi, "accepted_random_bits": acc_rnd,
                       "equals_cpu_scan": bool(np.array_equal(gi, ci) and np.array_equal(g1, c1) and acc_lo == int(ca.sum())),
                       "popc32_issued_per_s": 5 * 25e6 / (mms * 1e-3), "popc32_equivalent_per_s": 8 * 25e6 / (mms * 1e-3),
                       "popc32_peak_per_s": pk["popc32_per_s"] if pk else None,
                       "frac_of_popc_peak": (5 * 25e6 / (mms * 1e-3) / pk["popc32_per_s"]) if pk else None,
                       "frac_of_instruction_mix_peak": (25e6 / (mms * 1e-3) / pk["matcher_mix_pairs_per_s"]) if pk else None,
                       "roofline_note": "integer pipe, not HBM (320 KB of data): 5 POPC.32 + carry-save LOP3s per pair; peaks from scripts/ubench/popc on this GPU",
                       "cpu_baseline": {"ms_per_call": cpu_ms, "queries_per_s": 5000 / (cpu_ms * 1e-3), "cores": cores,
                                        "kind": "port (the scan of src/ORBmatcher.cc:574-605 with the reference's SWAR popcount, T threads over query rows)"}}
        except Exception as ex:                                            # the headline metric must still print
            matcher = {"error": repr(ex)}

    note("stages + matcher done")
    # ---- the "next" rows of SURVEY 8f, timed end to end through the C ABI (host wall clock, results on the host)
    extras = None
    if rank == 0 and not args.no_extras:
        extras = next_rows(orb, np, params, local, H, W)

    note("next rows done")
    # ---- CPU baselines beside it (rank 0, N=1 only), all host cores, bounded samples
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baselines(np, pool, params, affinity0)

    note("cpu baselines done")
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_dev / args.steps, "ms_per_step_per_rank": [round(v / args.steps, 4) for v in per_rank_dev], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": desc, "nfeatures": nfeat, "nlevels": nlev, "scale_factor": 1.2, "ini_th_fast": 20, "min_th_fast": 7,
                           "batch_per_gpu": batch, "global_batch": batch * world, "sharding": "frame-parallel, no collective", "cpu_bind": cpu_bind,
                           "dispatcher": "orbx_pool: %d native worker thread(s) per GPU, each batch split into %d contiguous sub-batch(es), %d batches in flight" % (lanes, lanes, depth),
                           "l2": "inputs larger than L2: %d frame slots = %.0f MB in HBM, walked cyclically" % (nslots, nslots * H * pitch / 2 ** 20),
                           "input_row_pitch": pitch, "keypoints_per_step": kp_dev / max(1, args.steps),
                           "fast_kernel": "k_fast_fused_tma (ORBX_OPT_FAST_TMA, experiment)" if args.fast_tma else "k_fast_fused (default)",
                           "result_records": ("orbx_keypoint_compact (12 B, lossless) + descriptor (32 B) per keypoint, %d B D2H per step" % (batch * (cap * 44 + 4))) if compact
                                             else "cv::KeyPoint (28 B) + descriptor (32 B) per keypoint, %d B D2H per step" % (batch * (cap * 60 + 4))},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": batch * H * W, "d2h_bytes_per_step": batch * (cap * 60 + 4),
                        "records": "cv::KeyPoint (28 B) + descriptor (32 B) per keypoint",
                        "ms_per_step": ms_e2e / args.steps, "ms_per_step_per_rank": [round(v / args.steps, 4) for v in per_rank_e2e]},
                "gpu_launches": launches, "parity_checked": bool(parity and parity["checked"]), "parity": parity,
                "roofline": roofline, "cpu_baseline": cpu, "matcher": matcher, "next_rows": extras}
        print(json.dumps(line))
    xpool.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baselines(np, pool, params, affinity0):
    """(1) the reference's own ORBextractor.cc compiled against the oracle's OpenCV-primitive shim, one extractor per
    thread, T = all host cores; (2) the Python + cv2 restatement as T single-threaded processes.  The faster one is "the" baseline."""
    from oracle.oracle import Oracle, RefExtractor
    if affinity0 is not None:
        os.sched_setaffinity(0, affinity0)                          # all host cores for the CPU legs
    cores = os.cpu_count() or 1
    chunk = np.ascontiguousarray(pool[:max(1, min(POOL_DISTINCT, cores))])
    if RefExtractor.available("asis"):
        kind, fn = "reference", lambda: RefExtractor.extract_many(params, chunk, cores, "asis")
    else:
        kind, fn = "port", lambda: Oracle.extract_many(params, chunk, cores)
    fn()
    tot_s, tot_f = 0.0, 0
    while tot_s < 10.0 and tot_f < 200000:
        s, _ = fn()
        tot_s += s; tot_f += len(chunk)
    legs = {kind: {"value": tot_f / tot_s, "unit": "frames/s", "cores": cores,
                   "sample": "%d frames of the same pool in chunks of %d, %.1f s wall, one extractor instance per thread" % (tot_f, len(chunk), tot_s),
                   "ms_per_frame_per_core": 1e3 * tot_s * min(cores, len(chunk)) / tot_f,
                   "what": "src/ORBextractor.cc of the reference compiled unmodified (oracle/_ref); OpenCV primitives = oracle/cvprim.c (AVX2 FAST / blur)" if kind == "reference"
                           else "oracle/orb_oracle.c"}}
    try:
        # a separate process: the workers are forked, which must not happen in a process that holds a CUDA context and helper threads
        H, W = pool.shape[1], pool.shape[2]
        procs = min(cores, len(chunk))
        out = subprocess.run([sys.executable, "-m", "oracle.cv2_restatement", str(H), str(W), str(params[0]), str(params[2]), str(len(chunk)), str(procs), "4"],
                             capture_output=True, text=True, timeout=240, cwd=ROOT)
        r = json.loads(out.stdout.strip().splitlines()[-1])
        legs["cv2"] = {"value": r["value"], "unit": "frames/s", "cores": procs,
                       "sample": "%d frames, %.1f s wall, %d single-threaded processes (cv2.setNumThreads(1))" % (r["frames"], r["wall_s"], procs),
                       "ms_per_frame_per_core": r["ms_per_frame_per_core"],
                       "what": "oracle/cv2_restatement.py: cv2 %s resize / FAST / GaussianBlur per the reference's call sites, octree in C, IC_Angle / rBRIEF in numpy" % r["cv2"]}
    except Exception as ex:
        legs["cv2"] = {"error": repr(ex)}
    best = max((k for k in legs if "value" in legs[k]), key=lambda k: legs[k]["value"])
    out = dict(legs[best])
    out["kind"] = "reference" if best == "reference" else ("port" if best in ("port", "cv2") else best)
    out["leg"] = best
    out["legs"] = legs
    return out


def next_rows(orb, np, params, local, H, W):
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
        return {"search_for_initialization": {"workload": "ORBmatcher::SearchForInitialization, %d x %d keypoints, window 100, ratio 0.9" % (len(kL), len(ic["xy2"])),
                                              "ms_per_call": 1e3 * t_init, "nmatches": int(ninit)},
                "search_by_projection": {"workload": "ORBmatcher::SearchByProjection(CurrentFrame, LastFrame), %d map points x %d features, th 15" % (len(kL), len(pc["cur_xy"])),
                                         "ms_per_call": 1e3 * t_proj, "nmatches": int(nproj)},
                "single_frame_call": {"workload": "ORBextractor::operator() on one %dx%d host frame (H2D, extract, D2H), the shape Frame::ExtractORB calls" % (W, H),
                                      "ms_per_frame": 1e3 * t_single},
                "stereo_match": {"workload": "Frame::ComputeStereoMatches, %d x %d keypoints, %dx%d pair" % (len(kL), len(kR), W, H),
                                 "ms_per_pair": 1e3 * t_stereo, "matches_kept": int(kept)},
                "bow_transform": {"workload": "ORBVocabulary::transform, %d descriptors, synthetic k=10 L=4 tree" % len(dL), "ms_per_frame": 1e3 * t_bow},
                "distinctive_descriptors": {"workload": "5000 map points, %d observations" % int(off[-1]), "ms_per_batch": 1e3 * t_dd},
                "note": "SURVEY 8f rows, host wall clock through the synchronous C-ABI calls, H2D / D2H of the small arrays included"}
    except Exception as ex:
        return {"error": repr(ex)}


if __name__ == "__main__":
    main()
