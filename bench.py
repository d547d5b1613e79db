#!/usr/bin/env python
"""bench.py -- lookahead frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one whole synthetic sequence (F frames of moving content) pushed through the lookahead:
every frame in through addPicture, every decided frame out through getDecidedPicture (+ the
per-frame results RateControl reads), then flush.  Work per step: F pre-lookaheads (K1-K3), every
motion search / frame cost the window needs (K4/K5), slice-type decisions, cuTree (K7/K8).

  value : frames/s with the pictures already resident in HBM when the timed region starts.
  e2e   : the same through the host API with pinned HOST pictures, H2D inside the timed region, and a
          D2H read of each decided frame's qp offsets + costs (what RateControl consumes).
  roofline : the dominant kernel (search_kernel, K4) against the measured HBM copy bandwidth, using the
          algorithmic bytes of SURVEY.md section 8d (5 lowres planes read + 12 B/block written per search).
  cpu_baseline / --impl reference : the UNMODIFIED reference lookahead (oracle/_ref, C primitives -- nasm is
          not in the image) with a thread pool over all host cores, on a bounded sample of the workload.

N > 1 (torchrun): every rank runs an independent stream on its own GPU (lookahead windows of separate
streams shard with no data-path collective); value = total frames / max-over-ranks step time; scaling "weak".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the >=20x target is quoted on
    "2160p-main10": dict(width=3840, height=2160, depth=10, frames=300, cpu_frames=72, seed=2,
                         la=dict(bframes=8, lookaheadDepth=60, bFrameAdaptive=2),
                         text="2160p main10 --rc-lookahead 60 --bframes 8 --b-adapt 2 cutree (BASELINE configs[1])"),
    # BASELINE.json configs[0]
    "1080p-8bit": dict(width=1920, height=1080, depth=8, frames=300, cpu_frames=100, seed=1,
                       la=dict(bframes=4, lookaheadDepth=20, bFrameAdaptive=2),
                       text="1080p 8-bit preset medium --rc-lookahead 20 --bframes 4 --b-adapt 2 (BASELINE configs[0])"),
    "360p-smoke": dict(width=640, height=360, depth=8, frames=60, cpu_frames=60, seed=1,
                       la=dict(bframes=4, lookaheadDepth=20, bFrameAdaptive=2), text="640x360 smoke"),
}


def gen_frames(wl, n, seed_offset=0):
    import _pkg
    synth = _pkg.load_synth()
    seq = synth.SynthSequence(wl["width"], wl["height"], depth=wl["depth"], seed=wl["seed"] + seed_offset,
                              cuts=(n // 2 + 3,), n_rects=6)
    return [seq.frame(i) for i in range(n)]


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, nm in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_reference_sample(wl, frames, threads):
    """frames/s of the unmodified reference lookahead (oracle/_ref) on `frames`, thread pool of `threads`."""
    import refbind
    if not refbind.available(wl["depth"]):
        return None
    ref = refbind.RefLookahead(wl["width"], wl["height"], depth=wl["depth"], poolThreads=threads, lookaheadSlices=0,
                               **wl["la"])
    t0 = time.perf_counter()
    for (y, u, v) in frames:
        ref.put(y, u, v, snap=False)
    ref.flush(snap=False)
    dt = time.perf_counter() - t0
    n_out = ref.lib.ref_la_num_out(ref.h)
    ref.close()
    assert n_out == len(frames), (n_out, len(frames))
    return len(frames) / dt, dt


def bench_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import refbind
    cores = os.cpu_count() or 1
    if not refbind.available(wl["depth"]):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built in this snapshot"}))
        return
    nfr = wl["cpu_frames"]
    frames = gen_frames(wl, nfr)
    times = []
    for i in range(args.warmup + args.steps):
        fps, dt = run_reference_sample(wl, frames, cores)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = nfr / (ms / 1000.0)
    line = {"impl": "reference", "metric": "lookahead_frames_per_s", "value": round(value, 3), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16" if wl["depth"] > 8 else "u8",
            "data": "synthetic", "config": {"workload": wl["text"], "frames_per_step": nfr, "lookahead_slices": 0},
            "cpu_baseline": {"value": round(value, 3), "unit": "frames/s", "cores": cores, "kind": "reference",
                             "sample": "first %d frames of the workload sequence; unmodified reference, C primitives "
                                       "(no nasm in the image => no asm), thread pool over all %d host cores" % (nfr, cores)},
            "e2e": {"value": round(value, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def secondary_workload(pkg, eng, name, args, device):
    """frames/s of another BASELINE config with the pictures resident in HBM (3 warm-up + 3 timed steps, CUDA-event
    timing through the engine's stopwatch), plus the reference lookahead on a sample of the same sequence."""
    import torch
    wl = WORKLOADS[name]
    F, depth, W, H = wl["frames"], wl["depth"], wl["width"], wl["height"]
    frames = gen_frames(wl, F)
    dev = [tuple(torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a).cuda() for a in f) for f in frames]
    torch.cuda.synchronize()
    kw = dict(wl["la"], asyncDepth=args.async_depth, speculate=args.speculate, pendingMax=args.pending_max or max(8, args.async_depth),
              batchMin=args.batch_min, device=device)
    times, types0 = [], None
    for it in range(6):
        la = pkg.Lookahead(W, H, depth=depth, **kw)
        ctx = la.engine()
        types, held = [], []
        eng.x265cu_sync(ctx)
        eng.x265cu_timer_start(ctx)

        def drain():
            while True:
                info = la.get_decided()
                if info is None:
                    return
                types.append(info.sliceType)
                held.append(info.handle)
                while len(held) > 2:
                    la.release(held.pop(0))
        for i, (y, u, v) in enumerate(dev):
            la.add_picture_ptr(y.data_ptr(), u.data_ptr(), v.data_ptr(), y.shape[1], u.shape[1], pts=i)
            drain()
        la.flush()
        drain()
        ms = C.c_double(0)
        eng.x265cu_timer_stop(ctx, C.byref(ms))
        la.close()
        assert len(types) == F
        assert types0 is None or types == types0
        types0 = types
        if it >= 3:
            times.append(ms.value)
    ms_step = float(np.mean(times))
    out = {"workload": wl["text"], "frames_per_step": F, "value": round(F / (ms_step / 1000.0), 2), "unit": "frames/s",
           "ms_per_step": round(ms_step, 3), "steps": 3, "warmup": 3, "dtype": "u16" if depth > 8 else "u8"}
    nfr = wl["cpu_frames"]
    r = run_reference_sample(wl, frames[:nfr], os.cpu_count() or 1)
    if r is not None:
        out["cpu_baseline"] = {"value": round(r[0], 3), "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "reference",
                               "sample": "first %d frames of the same sequence, %.1f s" % (nfr, r[1])}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2160p-main10", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--async-depth", type=int, default=32,
                    help="extra frames of input delay (LookaheadParam::asyncDepth): same decisions, GPU slack")
    ap.add_argument("--speculate", type=int, default=1)
    ap.add_argument("--shard", default="streams", choices=["streams", "window"],
                    help="N > 1: 'streams' = one independent stream per GPU (weak scaling, no data-path collective); "
                         "'window' = ONE stream whose searches / estimates are split over the GPUs by source frame, "
                         "stores exchanged by NCCL broadcast after every batch (strong scaling)")
    ap.add_argument("--pending-max", type=int, default=0, help="0 = async depth")
    ap.add_argument("--batch-min", type=int, default=0, help="LookaheadParam::batchMin (0 = async depth / 2)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.frames:
        wl["frames"] = args.frames

    if args.impl == "reference":
        bench_reference(args, wl)
        return

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the lookahead engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    import _pkg
    import importlib.util
    spec = importlib.util.spec_from_file_location("x265_amod_b200_shard", os.path.join(ROOT, "x265-amod_b200", "shard.py"))
    shard = importlib.util.module_from_spec(spec); spec.loader.exec_module(shard)
    dist = shard.init("nccl") if world > 1 else None
    pkg = _pkg.load_pkg()
    eng = pkg.load_engine()

    F = wl["frames"]
    depth, W, H = wl["depth"], wl["width"], wl["height"]
    window = args.shard == "window" and world > 1
    frames = gen_frames(wl, F, seed_offset=0 if window else rank)
    tdt = torch.uint8 if depth == 8 else torch.int16      # int16 views of the uint16 samples (bytes are what matter)

    def to_t(a):
        return torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)

    pinned_ok = [True]

    def pin(t):
        try:
            return t.pin_memory()
        except RuntimeError:        # page-locking refused (many ranks on one host): pageable uploads still work, slower
            pinned_ok[0] = False
            return t.clone()

    host = [(pin(to_t(y)), pin(to_t(u)), pin(to_t(v))) for (y, u, v) in frames]
    dev = [(y.cuda(), u.cuda(), v.cuda()) for (y, u, v) in host]
    # the pageable copies are only needed for the CPU baseline's sample (rank 0, N = 1): 7.5 GB per rank otherwise
    frames = frames[:wl["cpu_frames"]] if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    torch.cuda.synchronize()
    bytes_in = sum(t.numel() * t.element_size() for t in host[0])

    la_kw = dict(wl["la"], asyncDepth=args.async_depth, speculate=args.speculate,
                 pendingMax=args.pending_max or max(8, args.async_depth), batchMin=args.batch_min,
                 device=local_rank)
    exchange = None
    if window:
        la_kw["shardCount"] = world
        exchange = shard.make_exchange(dist, pkg.EXCHANGE_FN, cuda=True)
    geom = {}

    def one_step(pics, fetch_results):
        """returns (device ms, decided types, counters delta, profile)"""
        la = pkg.Lookahead(W, H, depth=depth, **la_kw)
        if window:
            la.shard(rank, world, exchange)
        g = la.geom
        geom.update(ncu=g.ncu, bw=g.bw, bh=g.bh, low_w=g.low_width, low_h=g.low_height)
        ctx = la.engine()
        ncu = g.ncu
        qp = np.zeros(ncu, np.float64); ic = np.zeros(ncu, np.int32); lc = np.zeros(ncu, np.uint16)
        fo = pkg.FrameOut()
        fo.qp_cutree_offset = qp.ctypes.data; fo.intra_cost = ic.ctypes.data; fo.lowres_costs00 = lc.ctypes.data
        d2h = [0]
        types = []
        last_nonb = [None, None]     # handles of the two most recent non-B frames (reference frames)
        pending = []

        def drain():
            while True:
                info = la.get_decided()
                if info is None:
                    return
                types.append(info.sliceType)
                if fetch_results:
                    # what RateControl / the frame encoder read per coded frame
                    if info.sliceType in (1, 2):
                        la.estimated_picture_cost(info.handle, None, None)
                    elif info.sliceType == 3 and last_nonb[1] is not None:
                        la.estimated_picture_cost(info.handle, last_nonb[1], None)
                    la.lib.x265la_frame_fetch(la.h, info.handle, C.byref(fo))
                    d2h[0] += qp.nbytes + ic.nbytes + lc.nbytes
                if info.sliceType in (1, 2, 3):
                    if last_nonb[0] is not None:
                        pending.append(last_nonb[0])
                    last_nonb[0], last_nonb[1] = last_nonb[1], info.handle
                    while pending:
                        la.release(pending.pop())
                else:
                    la.release(info.handle)

        cnt0 = pkg.Counters(); eng.x265cu_get_counters(ctx, C.byref(cnt0))
        eng.x265cu_profile_enable(ctx, 1)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        eng.x265cu_sync(ctx)
        eng.x265cu_timer_start(ctx)
        t0 = time.perf_counter()
        for i, (y, u, v) in enumerate(pics):
            la.add_picture_ptr(y.data_ptr(), u.data_ptr(), v.data_ptr(), y.shape[1], u.shape[1], pts=i)
            drain()
        la.flush()
        drain()
        ms = C.c_double(0)
        eng.x265cu_timer_stop(ctx, C.byref(ms))
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1000.0
        cnt1 = pkg.Counters(); eng.x265cu_get_counters(ctx, C.byref(cnt1))
        ht = (C.c_double * 10)()
        la.lib.x265la_get_timers(la.h, ht, 1)
        host_t = dict(prelookahead_wait=ht[0], weightp=ht[1], enqueue=ht[2], result_wait=ht[3], decisions=ht[4],
                      slicetype_decide=ht[5], calls=ht[6], add_picture_speculation=ht[7], estimated_picture_cost=ht[8],
                      fetch_mirrors=ht[9], wall=wall / 1000.0)
        pm = (C.c_double * 7)(); pn = (C.c_uint64 * 7)(); pb = (C.c_double * 7)()
        eng.x265cu_profile_get_busy(ctx, pb)
        eng.x265cu_profile_get(ctx, pm, pn, 1)
        prof = {k: (pm[i], int(pn[i]), pb[i]) for i, k in enumerate(pkg.K_NAMES)}
        delta = dict(launches=cnt1.kernel_launches - cnt0.kernel_launches, h2d=cnt1.h2d_bytes - cnt0.h2d_bytes,
                     d2h=(cnt1.d2h_bytes - cnt0.d2h_bytes), search_jobs=cnt1.search_jobs - cnt0.search_jobs,
                     cost_jobs=cnt1.cost_jobs - cnt0.cost_jobs)
        assert len(types) == len(pics), (len(types), len(pics))
        la.close()
        prof["host"] = host_t
        return max(ms.value, 0.0), wall, types, delta, prof

    def reduce_max(x):
        return shard.reduce_max(dist, x, device="cuda")

    # --- value: pictures resident in HBM ---------------------------------------------------------------
    for _ in range(args.warmup):
        one_step(dev, False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    times, profs, deltas, types0 = [], [], [], None
    for _ in range(args.steps):
        ms, wall, types, delta, prof = one_step(dev, False)
        times.append(reduce_max(ms)); profs.append(prof); deltas.append(delta); types0 = types
    # --- e2e: pinned host pictures, H2D + result D2H inside the timed region ----------------------------
    e2e_times, e2e_delta, e2e_prof = [], None, None
    one_step(host, True)
    for _ in range(args.steps):
        ms, wall, types, delta, prof = one_step(host, True)
        e2e_times.append(reduce_max(ms)); e2e_delta = delta; e2e_prof = prof
        assert types == types0, "decisions differ between device-resident and host-fed runs"
    clocks = sampler.stop()

    ms_step = float(np.mean(times))
    n_streams = 1 if window else world
    value = n_streams * F / (ms_step / 1000.0)
    e2e_ms = float(np.mean(e2e_times))
    e2e_value = n_streams * F / (e2e_ms / 1000.0)

    # --- roofline of the dominant kernel (K4, motion search) ---------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    bpp = 2 if depth > 8 else 1
    P = geom["low_w"] * geom["low_h"] * bpp
    bytes_per_search = 5 * P + 12 * geom["ncu"]            # fenc P + ref 4P read, 12 B/block written
    # batches overlap on the GPU: the denominator is the time during which at least one search launch was running
    # (union of the CUDA-event intervals of the launches), not the sum of the individual launch durations
    s_ms = sum(p["search"][2] for p in profs); s_launch = sum(p["search"][1] for p in profs)
    s_sum_ms = sum(p["search"][0] for p in profs)
    s_jobs = sum(d["search_jobs"] for d in deltas)
    achieved = (s_jobs * bytes_per_search / 1e9) / (s_ms / 1000.0) if s_ms > 0 else 0.0
    kernel_ms = {k: round(sum(p[k][2] for p in profs) / len(profs), 3) for k in pkg.K_NAMES}
    # dram__bytes_read + dram__bytes_write of the kernel from the committed ncu --set full capture, scaled from that
    # capture's job count to this run's average launch (same unit as `achieved`: per launch)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_traffic.json"))).get(args.workload)
        if tj and s_launch:
            traffic = int(tj["dram_bytes_per_search_job"] * s_jobs / s_launch)
            traffic_src = tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "search_kernel (K4 motion search)", "achieved": round(achieved, 2), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 5), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": int(bytes_per_search * s_jobs / max(1, s_launch)), "peak_source": peak_src,
                "algorithmic_bytes_per_search_job": bytes_per_search, "search_jobs_per_step": s_jobs // max(1, len(deltas)),
                "search_launches_per_step": s_launch // max(1, len(profs)),
                "avg_launch_ms": round(s_sum_ms / max(1, s_launch), 4),
                "search_busy_ms_per_step": round(s_ms / max(1, len(profs)), 3),
                "kernel_busy_ms_per_step": kernel_ms,
                "host_ms_per_step": {k: round(1000.0 * sum(p["host"][k] for p in profs) / len(profs), 2) for k in profs[0]["host"]},
                "note": "K4 runs out of L2 and is bound by the ALU pipe / wavefront latency, not HBM (SURVEY 8d; ncu at "
                        "full occupancy: ALU pipe 65 %, issue 57 %, DRAM 0.7 %); the HBM fraction is the conservative "
                        "checkable figure, profiles/ holds the pipe utilisation"}

    line = {"metric": "lookahead_frames_per_s", "value": round(value, 2), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "higher_is_better": True,
            "scaling": "strong" if window else "weak", "vs_baseline": None, "dtype": "u16" if depth > 8 else "u8", "data": "synthetic",
            "config": {"workload": wl["text"], "frames_per_step": F, "resolution": "%dx%d" % (W, H), "bit_depth": depth,
                       "lookahead_slices": 0, "async_depth": args.async_depth, "speculate": args.speculate, "streams": n_streams,
                       "parallelism": ("one stream, searches/estimates split by source frame over %d GPUs, NCCL broadcast of the "
                                       "stores per batch" % world) if window else
                                      ("independent stream per GPU" if world > 1 else "1 GPU"),
                       "l2_policy": "inputs larger than L2: %d MB of pictures per step vs 126 MB L2" % (F * bytes_in // (1 << 20))},
            "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "ms_per_step": round(e2e_ms, 3),
                    "h2d_bytes_per_step": int(e2e_delta["h2d"]), "d2h_bytes_per_step": int(e2e_delta["d2h"]),
                    "host_memory": "pinned" if pinned_ok[0] else "pageable (page-locking was refused)",
                    "host_ms_last_step": {k: round(1000.0 * v, 2) for k, v in e2e_prof["host"].items()}},
            "gpu_launches": int(sum(d["launches"] for d in deltas)),
            "clocks": clocks, "roofline": roofline,
            "decided_types": "".join(pkg.TYPE_NAMES[t][0] if t != 4 else "b" for t in types0[:48])}

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "2160p-main10" and not args.frames:
        # the metric also names 1080p: BASELINE configs[0] as a secondary, shorter measurement (resident pictures, the
        # same timing rules) next to the reference on the same sequence.  Never allowed to break the main line.
        try:
            line["other_workloads"] = {"1080p-8bit": secondary_workload(pkg, eng, "1080p-8bit", args, local_rank)}
        except Exception as e:      # pragma: no cover
            line["other_workloads"] = {"1080p-8bit": {"error": repr(e)}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nfr = wl["cpu_frames"]
        r = run_reference_sample(wl, frames[:nfr], cores)
        if r is not None:
            line["cpu_baseline"] = {"value": round(r[0], 3), "unit": "frames/s", "cores": cores, "kind": "reference",
                                    "sample": "first %d frames of the same sequence, %.1f s; unmodified reference lookahead, C primitives "
                                              "(no nasm => no asm), thread pool over all %d host cores" % (nfr, r[1], cores)}
        else:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference",
                                    "sample": "oracle/_ref not present in this snapshot"}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
