#!/usr/bin/env python
"""bench.py -- lookahead frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

A "step" is one whole synthetic sequence (F frames of moving content) pushed through the lookahead:
every frame in through addPicture, every decided frame out through getDecidedPicture (+ the
per-frame results the encoder reads), then flush.  Work per step: F pre-lookaheads (K1-K3), every
motion search / frame cost the window needs (K4/K5), slice-type decisions, cuTree (K7/K8).

  value : frames/s with the pictures already resident in HBM when the timed region starts.
  e2e   : the same through the host API with pinned HOST pictures, H2D inside the timed region, and the D2H
          mirror of everything the main encoder reads of a decided frame (getEstimatedPictureCost, qp offsets,
          intra / block costs of the coded estimate, lowres MVs of the coded references, the four lowres planes
          of every non-B frame for weightPrediction) -- exactly what the ENABLE_CUDA build of the reference
          (integration/) copies into Frame::m_lowres.
  roofline : the dominant kernel (search_kernel, K4) against the measured HBM copy bandwidth, using the
          algorithmic bytes of SURVEY.md section 8d (5 lowres planes read + 12 B/block written per search).
  parity_checked : the first `cpu_frames` frames of the very sequence that is timed, same configuration, run through the
          CUDA lookahead and through the live unmodified reference, every published Lowres field compared
          (tests/compare.py); the run fails on a mismatch.
  cpu_baseline / --impl reference : the reference lookahead (oracle/_ref) with a thread pool over all host cores on a
          bounded sample of the workload, with the lookahead's hot primitives replaced by SSE4.1 intrinsics shims
          (oracle/ref_simd.cpp, verified bit-exact; nasm is not in the image, so this is the asm-CLASS denominator the
          metric asks for) and, next to it, with the plain C primitives.

Both arms run the same configuration, pool size included: x265 sizes its pool to the host cores and its lookahead
results depend on that size (slicetype.cpp:1024,2691,2733), so the CUDA arm emulates a pool of the same size.

N > 1 (torchrun): every rank runs an independent stream on its own GPU (lookahead windows of separate
streams shard with no data-path collective); value = total frames / max-over-ranks step time; scaling "weak".
"""
import argparse
import os
# 22 CUDA streams per lookahead context (16 batch lanes, 3 pre-lookahead, main, mirror, copy): give them their own hardware
# queues instead of the default 8, or unrelated streams falsely serialise behind each other's event waits
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the >=20x target is quoted on
    "2160p-main10": dict(width=3840, height=2160, depth=10, frames=300, cpu_frames=72, seed=2,
                         la=dict(bframes=8, lookaheadDepth=60, bFrameAdaptive=2),
                         text="2160p main10 --rc-lookahead 60 --bframes 8 --b-adapt 2 cutree (BASELINE configs[1])"),
    # BASELINE.json configs[0]
    "1080p-8bit": dict(width=1920, height=1080, depth=8, frames=300, cpu_frames=100, seed=1,
                       la=dict(bframes=4, lookaheadDepth=20, bFrameAdaptive=2),
                       text="1080p 8-bit preset medium --rc-lookahead 20 --bframes 4 --b-adapt 2 (BASELINE configs[0])"),
    # BASELINE.json configs[2]: preset slower's lookahead depths, weightp on, scenecut-heavy content (seeded cuts every
    # 12-40 frames, two 30-frame fades, two flashes; --me star --subme 5 only change the main encoder, the lookahead's
    # motion search is always HEX with subpel refine 1, slicetype.cpp:4114-4120)
    "1080p-slower-weightp": dict(width=1920, height=1080, depth=8, frames=300, cpu_frames=100, seed=3,
                                 la=dict(bframes=8, lookaheadDepth=40, bFrameAdaptive=2, bEnableWeightedPred=1),
                                 content="scenecut-heavy",
                                 text="1080p 8-bit preset slower (--rc-lookahead 40 --bframes 8) --weightp, scenecut-heavy content "
                                      "(BASELINE configs[2])"),
    # BASELINE.json configs[3] on ONE GPU (N > 1 adds the same stream sharded over the GPUs)
    "4320p-8bit": dict(width=7680, height=4320, depth=8, frames=160, cpu_frames=24, seed=4,
                       la=dict(bframes=4, lookaheadDepth=80, bFrameAdaptive=2),
                       text="4320p 8-bit --rc-lookahead 80 --bframes 4 (BASELINE configs[3])"),
    # --hme on BASELINE configs[0]: levels 0 (hex) and 1 (umh) of the hierarchical search on the GPU (a correctness-first
    # path -- the four blocks of a warp search independently); measured so that the cost of the option is on record
    "1080p-8bit-hme": dict(width=1920, height=1080, depth=8, frames=120, cpu_frames=40, seed=1,
                           la=dict(bframes=4, lookaheadDepth=20, bFrameAdaptive=2, bEnableHME=1),
                           text="1080p 8-bit preset medium --rc-lookahead 20 --bframes 4 --b-adapt 2 --hme (hex, umh; ranges 16, 32)"),
    "360p-smoke": dict(width=640, height=360, depth=8, frames=60, cpu_frames=60, seed=1,
                       la=dict(bframes=4, lookaheadDepth=20, bFrameAdaptive=2), text="640x360 smoke"),
}


def make_seq(wl, n, seed_offset=0):
    import _pkg
    synth = _pkg.load_synth()
    seed = wl["seed"] + seed_offset
    if wl.get("content") == "scenecut-heavy":
        rng = np.random.default_rng(1000 + seed)
        cuts, t = [], 0
        while True:
            t += int(rng.integers(12, 41))
            if t >= n:
                break
            cuts.append(t)
        fades = [(n // 5, 30, 0.3), (3 * n // 5, 30, 1.0)] if n >= 150 else [(n // 5, 10, 0.3)]
        flashes = [(2 * n // 5 + 2, 1), (4 * n // 5 + 1, 2)]
        return synth.SynthSequence(wl["width"], wl["height"], depth=wl["depth"], seed=seed, cuts=tuple(cuts), fades=fades,
                                   flashes=flashes, n_rects=6)
    return synth.SynthSequence(wl["width"], wl["height"], depth=wl["depth"], seed=seed, cuts=(n // 2 + 3,), n_rects=6)


def gen_frames(wl, n, seed_offset=0):
    seq = make_seq(wl, n, seed_offset)
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


# ------------------------------------------------------------------------------------------------ reference (CPU) arm

def run_reference_sample(wl, frames, threads, simd=True, snap=False):
    """frames/s of the reference lookahead (oracle/_ref) on `frames` with a thread pool of `threads`.  simd: install the
    SSE4.1 intrinsics shims for the hot primitives (asm-class baseline) instead of the C primitives.  snap: keep every
    decided frame's Lowres state inside the harness (for the parity check); the snapshot time is taken off the clock.
    Returns (fps, seconds, handle or None)."""
    import refbind
    if not refbind.available(wl["depth"]):
        return None
    ref = refbind.RefLookahead(wl["width"], wl["height"], depth=wl["depth"], poolThreads=threads, lookaheadSlices=0,
                               **ref_kwargs(wl["la"]))
    active = refbind.simd_install(wl["depth"], simd)
    try:
        t0 = time.perf_counter()
        for (y, u, v) in frames:
            ref.put(y, u, v, snap=snap)
        ref.flush(snap=snap)
        dt = time.perf_counter() - t0 - ref.snapshot_seconds()
    finally:
        refbind.simd_install(wl["depth"], False)
    assert ref.num_out() == len(frames), (ref.num_out(), len(frames))
    if not snap:
        ref.close()
        ref = None
    return len(frames) / dt, dt, ref, active


def ref_kwargs(la):
    """our LaParam names -> the reference harness' (oracle/refbind.py)"""
    m = dict(bEnableWeightedPred="weightp", bEnableWeightedBiPred="weightb", bEnableHME="hme")
    return {m.get(k, k): v for k, v in la.items()}


def primitives_text(active):
    return ("hot primitives = SSE4.1 intrinsics shims verified bit-exact against the C ones (oracle/ref_simd.cpp): asm-class"
            if active else "C primitives (no nasm in the image => no asm)")


def bench_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import refbind
    cores = os.cpu_count() or 1
    if not refbind.available(wl["depth"]):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not built in this snapshot"}))
        return
    simd = not args.c_primitives
    if simd and refbind.simd_selftest(wl["depth"], 6000) != 0:
        simd = False
    # a bounded sample per step: the whole sequence when the run has few steps, else its first cpu_frames frames
    nfr = wl["frames"] if (args.steps + args.warmup) <= 6 else wl["cpu_frames"]
    frames = gen_frames(wl, nfr)
    times, active = [], False
    for i in range(args.warmup + args.steps):
        fps, dt, _, active = run_reference_sample(wl, frames, cores, simd=simd)
        if i >= args.warmup:
            times.append(dt)
    ms = 1000.0 * float(np.mean(times))
    value = nfr / (ms / 1000.0)
    line = {"impl": "reference", "metric": "lookahead_frames_per_s", "value": round(value, 3), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16" if wl["depth"] > 8 else "u8",
            "data": "synthetic", "config": {"workload": wl["text"], "frames_per_step": nfr, "lookahead_slices": 0, "pool_workers": cores},
            "cpu_baseline": {"value": round(value, 3), "unit": "frames/s", "cores": cores,
                             "kind": "reference+intrinsics" if active else "reference",
                             "sample": "first %d frames of the workload sequence; reference lookahead, thread pool over all %d host "
                                       "cores; %s" % (nfr, cores, primitives_text(active))},
            "e2e": {"value": round(value, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ CUDA arm

class Stream:
    """one Lookahead context and the pictures it is fed"""

    def __init__(self, pkg, wl, pics, la_kw, mirror):
        self.pkg, self.wl, self.pics, self.la_kw, self.mirror = pkg, wl, pics, la_kw, mirror
        self.la = None

    def open(self, shard=None):
        wl = self.wl
        self.la = self.pkg.Lookahead(wl["width"], wl["height"], depth=wl["depth"], **self.la_kw)
        if shard:
            self.la.shard(*shard)
        g = self.la.geom
        self.geom = dict(ncu=g.ncu, bw=g.bw, bh=g.bh, low_w=g.low_width, low_h=g.low_height, ncu_full=g.ncu_full,
                         stride=g.stride, plane_lines=g.plane_lines)
        sm = (C.c_int32 * 2)()
        self.pkg.load_engine().x265cu_sm_partition(C.c_void_p(self.la.engine()), C.byref(sm, 0), C.byref(sm, 4))
        self.geom["sm_partition"] = [sm[0], sm[1]]
        self.types, self.d2h = [], 0
        self.tracker = self.pkg.RefTracker()
        self.pocs = {}
        self.fed = 0
        if self.mirror:
            # two sets of page-locked destination buffers: the mirror of frame i lands while frame i + 1 is decided
            ncu, nf, bpp, nb = g.ncu, g.ncu_full, 2 if wl["depth"] > 8 else 1, g.nb
            self.sets, self.tickets, self.nmir = [], [None] * 4, 0
            for _ in range(4):
                b = dict(qpAq=np.zeros(nf, np.float64), qpCt=np.zeros(nf, np.float64), invQ=np.zeros(nf, np.int32),
                         intra=np.zeros(ncu, np.int32), lc=np.zeros(ncu, np.uint16), rs=np.zeros(g.bh, np.int32),
                         mv=np.zeros((2, nb, ncu, 2), np.int32),
                         planes=np.zeros(4 * g.stride * g.plane_lines, np.uint16 if bpp == 2 else np.uint8))
                for a in b.values():
                    self.la.lib.x265la_pin(self.la.h, a.ctypes.data, a.nbytes)
                m = self.pkg.Mirror()
                m.intraCost = b["intra"].ctypes.data; m.qpAqOffset = b["qpAq"].ctypes.data; m.qpCuTreeOffset = b["qpCt"].ctypes.data
                m.invQscaleFactor = b["invQ"].ctypes.data; m.lowresCosts = b["lc"].ctypes.data; m.rowSatds = b["rs"].ctypes.data
                for l in range(2):
                    for d in range(1, nb):
                        m.lowresMvs[l][d] = b["mv"][l, d].ctypes.data
                self.sets.append((b, m))
        self.weightp = bool(self.la.param.bEnableWeightedPred)
        self.pub = (C.c_uint32 * 2)()

    def feed(self):
        if self.fed >= len(self.pics):
            return False
        y, u, v = self.pics[self.fed]
        self.la.add_picture_ptr(y.data_ptr(), u.data_ptr(), v.data_ptr(), y.shape[1], u.shape[1], pts=self.fed)
        self.fed += 1
        return True

    def drain(self, limit=1 << 30):
        """take up to `limit` decided frames (Encoder::encode takes one per picture it adds, encoder.cpp:2130)"""
        la = self.la
        for _ in range(limit):
            info = la.get_decided()
            if info is None:
                return
            self.types.append(info.sliceType)
            self.pocs[info.handle] = info.poc
            old = set(t for _, t in self.tracker.refs)
            r0, r1 = self.tracker.push(info.poc, info.sliceType, info.handle)
            if self.mirror:
                # what Encoder::encode / RateControl / the frame encoder / search / weightPrediction read of a decided frame
                # (SURVEY 8b "output contract"): the same requests the ENABLE_CUDA Lookahead shim makes (integration/)
                k = self.nmir & 3
                b, m = self.sets[k]
                if self.tickets[k] is not None:
                    la.lib.x265la_mirror_wait(la.h, self.tickets[k])
                d0 = info.poc - self.pocs[r0] if r0 else 0
                d1 = self.pocs[r1] - info.poc if r1 else 0
                la.lib.x265la_estimated_picture_cost_dist(la.h, info.handle, d0, d1)
                m.d0, m.d1 = d0, d1
                isb = info.sliceType in (self.pkg.TYPE_B, self.pkg.TYPE_BREF)
                m.planes = b["planes"].ctypes.data if (self.weightp and not isb) else None
                t = C.c_int64(0)
                if la.lib.x265la_frame_mirror_async(la.h, info.handle, C.byref(m), self.pub, C.byref(t)) != 0:
                    raise RuntimeError("mirror failed: %s" % la.lib.x265la_last_error(la.h).decode())
                self.tickets[k] = t.value
                self.nmir += 1
            live = set(t for _, t in self.tracker.refs)
            for h in old - live:
                la.release(h)
            if info.handle not in live:
                la.release(info.handle)

    def finish(self):
        if self.mirror:
            for t in self.tickets:
                if t is not None:
                    self.la.lib.x265la_mirror_wait(self.la.h, t)

    def close(self):
        if self.mirror:
            for b, _ in self.sets:
                for a in b.values():
                    self.la.lib.x265la_unpin(self.la.h, a.ctypes.data)
        self.la.close()
        self.la = None


def run_step(eng, streams, dist=None, shard=None, profile=True):
    """one step over `streams` (fed round-robin, one picture each per turn).  Returns the device time (ms, CUDA events: the
    slowest context's stopwatch, all started after a common synchronisation), and per stream (types, counters, profile)."""
    import torch
    for s in streams:
        s.open(shard)
    ctxs = [s.la.engine() for s in streams]
    cnt0 = []
    for c in ctxs:
        k = streams[0].pkg.Counters(); eng.x265cu_get_counters(c, C.byref(k)); cnt0.append(k)
        if profile:
            eng.x265cu_profile_enable(c, 1)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    for c in ctxs:
        eng.x265cu_sync(c)
    for c in ctxs:
        eng.x265cu_timer_start(c)
    t0 = time.perf_counter()
    more = True
    while more:
        more = False
        for s in streams:
            if s.feed():
                more = True
                s.drain(1)
    for s in streams:
        s.la.flush()
        s.drain()
        s.finish()
    ms = 0.0
    for c in ctxs:
        m = C.c_double(0)
        eng.x265cu_timer_stop(c, C.byref(m))
        ms = max(ms, m.value)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1000.0
    out = []
    for s, c, k0 in zip(streams, ctxs, cnt0):
        pkg = s.pkg
        k1 = pkg.Counters(); eng.x265cu_get_counters(c, C.byref(k1))
        ht = (C.c_double * 10)()
        s.la.lib.x265la_get_timers(s.la.h, ht, 1)
        host_t = dict(prelookahead_wait=ht[0], weightp=ht[1], enqueue=ht[2], result_wait=ht[3], decisions=ht[4],
                      slicetype_decide=ht[5], calls=ht[6], add_picture_speculation=ht[7], estimated_picture_cost=ht[8],
                      fetch_mirrors=ht[9], wall=wall / 1000.0)
        prof = {}
        if profile:
            pm = (C.c_double * 7)(); pn = (C.c_uint64 * 7)(); pb = (C.c_double * 7)()
            eng.x265cu_profile_get_busy(c, pb)
            eng.x265cu_profile_get(c, pm, pn, 1)
            prof = {k: (pm[i], int(pn[i]), pb[i]) for i, k in enumerate(pkg.K_NAMES)}
        prof["host"] = host_t
        delta = dict(launches=k1.kernel_launches - k0.kernel_launches, h2d=k1.h2d_bytes - k0.h2d_bytes,
                     d2h=k1.d2h_bytes - k0.d2h_bytes, search_jobs=k1.search_jobs - k0.search_jobs,
                     cost_jobs=k1.cost_jobs - k0.cost_jobs)
        assert len(s.types) == len(s.pics), (len(s.types), len(s.pics))
        out.append((list(s.types), delta, prof, dict(s.geom)))
        s.close()
    return max(ms, 0.0), wall, out


def to_t(a):
    import torch
    return torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else a)


def parity_check(pkg, wl, frames, ref, la_kw):
    """The first len(frames) frames of the timed sequence through the CUDA lookahead (same configuration, results fetched
    frame by frame) against the reference's snapshots held by `ref`.  Returns the parity block of the JSON line."""
    import compare
    la = pkg.Lookahead(wl["width"], wl["height"], depth=wl["depth"], **la_kw)
    bad, idx = [], [0]
    rkw = wl["la"]

    def drain():
        while True:
            info = la.get_decided()
            if info is None:
                return
            got = la.frame_dict(info, planes=False)
            la.release(info.handle)
            want = ref.frame(idx[0], drop=True)
            idx[0] += 1
            if len(bad) < 12:
                bad.extend(compare.compare_frames(want, got, cutree=rkw.get("cuTree", 1), weightp=rkw.get("bEnableWeightedPred", 1)))
    for i, (y, u, v) in enumerate(frames):
        la.add_picture(y, u, v, pts=i)
        drain()
    la.flush()
    drain()
    la.close()
    if idx[0] != len(frames):
        bad.append("frame count: ours %d, input %d" % (idx[0], len(frames)))
    return {"frames": idx[0], "mismatches": len(bad), "against": "live unmodified reference (oracle/_ref), same sequence, same configuration",
            "fields": "sliceType bScenecut bKeyframe leadingBframes costEst costEstAq intraMbs lowresMvs lowresMvCosts lowresCosts rowSatds "
                      "intraCost intraMode invQscaleFactor wp_ssd wp_sum propagateCost (exact); qpAqOffset qpCuTreeOffset (<= 1e-3 QP)",
            "first_mismatches": bad[:4]}


def measure_workload(pkg, eng, wl, args, device, cores, steps, warmup, with_e2e, with_parity, with_cpu, seed_offset=0, frames=None):
    """value (+ e2e, parity, cpu baseline) of one single-stream workload on this rank's GPU"""
    import torch
    F = wl["frames"]
    if frames is None:
        frames = gen_frames(wl, F, seed_offset)
    pinned_ok = [True]

    def pin(t):
        try:
            return t.pin_memory()
        except RuntimeError:        # page-locking refused (many ranks on one host): pageable uploads still work, slower
            pinned_ok[0] = False
            return t.clone()
    la_kw = dict(wl["la"], asyncDepth=args.async_depth, speculate=args.speculate, pendingMax=args.pending_max or max(8, args.async_depth),
                 batchMin=args.batch_min, device=device, poolWorkers=cores)
    dev = [tuple(to_t(a).cuda() for a in f) for f in frames]
    torch.cuda.synchronize()
    res = {"la_kw": la_kw}
    for _ in range(warmup):
        run_step(eng, [Stream(pkg, wl, dev, la_kw, False)])
    times, profs, deltas, types0, geom = [], [], [], None, None
    for _ in range(steps):
        ms, wall, out = run_step(eng, [Stream(pkg, wl, dev, la_kw, False)])
        types, delta, prof, geom = out[0]
        assert types0 is None or types == types0, "decisions differ between two runs of the same sequence"
        times.append(ms); profs.append(prof); deltas.append(delta); types0 = types
    res.update(times=times, profs=profs, deltas=deltas, types=types0, geom=geom)
    if with_e2e:
        host = [tuple(pin(to_t(a)) for a in f) for f in frames]
        e2e_kw = dict(la_kw, extraSlots=12)
        run_step(eng, [Stream(pkg, wl, host, e2e_kw, True)])
        e_times = []
        for _ in range(steps):
            ms, wall, out = run_step(eng, [Stream(pkg, wl, host, e2e_kw, True)])
            types, delta, prof, _ = out[0]
            assert types == types0, "decisions differ between device-resident and host-fed runs"
            e_times.append(ms)
        res.update(e2e_times=e_times, e2e_delta=delta, e2e_prof=prof, pinned=pinned_ok[0],
                   bytes_in=sum(t.numel() * t.element_size() for t in host[0]))
        del host
    del dev
    torch.cuda.empty_cache()
    if with_cpu:
        nfr = min(wl["cpu_frames"], F)
        sample = frames[:nfr]
        import refbind
        ok = refbind.available(wl["depth"]) and refbind.simd_selftest(wl["depth"], 6000) == 0
        r = run_reference_sample(wl, sample, cores, simd=ok, snap=with_parity)
        if r is not None:
            fps, dt, ref, active = r
            if with_parity:
                res["parity"] = parity_check(pkg, wl, sample, ref, dict(la_kw, asyncDepth=min(16, args.async_depth)))
                ref.close()
            rc = run_reference_sample(wl, sample, cores, simd=False) if active else None
            res["cpu_baseline"] = {"value": round(fps, 3), "unit": "frames/s", "cores": cores,
                                   "kind": "reference+intrinsics" if active else "reference",
                                   "sample": "first %d frames of the same sequence, %.1f s; reference lookahead, thread pool over all %d "
                                             "host cores; %s" % (nfr, dt, cores, primitives_text(active))}
            if rc is not None:
                res["cpu_baseline"]["c_primitives_value"] = round(rc[0], 3)
                res["cpu_baseline"]["c_primitives_note"] = "the same sample with the reference's plain C primitives (its no-asm build), %.1f s" % rc[1]
        else:
            res["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference",
                                   "sample": "oracle/_ref not present in this snapshot"}
    return res


def secondary_line(wl, res, F):
    ms = float(np.mean(res["times"]))
    out = {"workload": wl["text"], "frames_per_step": F, "value": round(F / (ms / 1000.0), 2), "unit": "frames/s",
           "ms_per_step": round(ms, 3), "steps": len(res["times"]), "dtype": "u16" if wl["depth"] > 8 else "u8",
           "pool_workers": res["la_kw"]["poolWorkers"],
           "decided_types": "".join("?IIPbB"[t] if t != 1 else "I" for t in res["types"][:40])}
    if "e2e_times" in res:
        e = float(np.mean(res["e2e_times"]))
        out["e2e"] = {"value": round(F / (e / 1000.0), 2), "unit": "frames/s", "ms_per_step": round(e, 3),
                      "h2d_bytes_per_step": int(res["e2e_delta"]["h2d"]), "d2h_bytes_per_step": int(res["e2e_delta"]["d2h"])}
    for k in ("parity", "cpu_baseline"):
        if k in res:
            out[k] = res[k]
    return out


def box_down(a, k):
    """k x k box downscale with rounding (abrEncApp-style rendition source)"""
    h, w = a.shape[0] // k * k, a.shape[1] // k * k
    s = a[:h, :w].astype(np.uint32).reshape(h // k, k, w // k, k).sum(axis=(1, 3))
    return ((s + (k * k) // 2) // (k * k)).astype(a.dtype)


def ladder_workload(pkg, eng, args, device, cores, src_frames, rank, world, dist, reduce_max):
    """BASELINE configs[4]: the 2160p / 1080p / 720p rendition ladder, one independent lookahead per rendition (abrEncApp
    runs one Lookahead per rendition, abrEncApp.cpp:510).  Renditions are dealt round-robin over the ranks: three concurrent
    contexts on one GPU at N = 1, one GPU each at N >= 3.  A ladder frame is done when every rendition has decided it."""
    import torch
    base = WORKLOADS["2160p-main10"]
    n = len(src_frames)
    rend = []
    for name, k in (("2160p", 1), ("1080p", 2), ("720p", 3)):
        wl = dict(base, width=3840 // k, height=2160 // k, frames=n, text="ladder rendition %s main10" % name)
        rend.append((name, k, wl))
    mine = [r for i, r in enumerate(rend) if i % world == rank]
    streams = []
    for name, k, wl in mine:
        fr = src_frames if k == 1 else [(box_down(y, k), box_down(u, k), box_down(v, k)) for (y, u, v) in src_frames]
        dev = [tuple(to_t(a).cuda() for a in f) for f in fr]
        la_kw = dict(wl["la"], asyncDepth=args.async_depth, speculate=args.speculate, pendingMax=max(8, args.async_depth),
                     batchMin=args.batch_min, device=device, poolWorkers=cores)
        streams.append((wl, dev, la_kw))
    torch.cuda.synchronize()
    times = []
    for it in range(5):
        if streams:
            ms, wall, out = run_step(eng, [Stream(pkg, wl, dev, kw, False) for (wl, dev, kw) in streams], dist=dist, profile=False)
        else:
            if dist is not None:
                dist.barrier()
            ms = 0.0
        ms = reduce_max(ms)
        if it >= 2:
            times.append(ms)
    ms = float(np.mean(times))
    return {"workload": "2160p / 1080p / 720p main10 rendition ladder, --rc-lookahead 60 --bframes 8 per rendition (BASELINE configs[4])",
            "frames_per_step": n, "renditions": [r[0] for r in rend],
            "placement": "%d rendition context(s) per GPU over %d GPU(s)" % ((len(rend) + world - 1) // world, min(world, len(rend))),
            "value": round(n / (ms / 1000.0), 2), "unit": "ladder frames/s (all renditions of a source frame decided)",
            "rendition_frames_per_s": round(3 * n / (ms / 1000.0), 2), "ms_per_step": round(ms, 3), "steps": 3, "warmup": 2}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="2160p-main10", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the end-to-end measurement (the line then repeats `value` there and says so)")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary workloads (other BASELINE configs)")
    ap.add_argument("--c-primitives", action="store_true", help="--impl reference: time the plain C primitives instead of the intrinsics shims")
    ap.add_argument("--async-depth", type=int, default=64,
                    help="extra frames of input delay (LookaheadParam::asyncDepth): same decisions, GPU slack")
    ap.add_argument("--speculate", type=int, default=1)
    ap.add_argument("--shard", default="streams", choices=["streams", "window"],
                    help="N > 1: 'streams' = one independent stream per GPU (weak scaling, no data-path collective); "
                         "'window' = ONE stream whose searches / estimates are split over the GPUs by source frame "
                         "(strong scaling)")
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
    cores = os.cpu_count() or 1

    def reduce_max(x):
        return shard.reduce_max(dist, x, device="cuda")

    F = wl["frames"]
    depth, W, H = wl["depth"], wl["width"], wl["height"]
    window = args.shard == "window" and world > 1
    main_rank = rank == 0 and world == 1
    t_start = time.perf_counter()
    frames = gen_frames(wl, F, seed_offset=0 if window else rank)

    if window:
        line = window_shard_line(pkg, eng, shard, dist, wl, frames, args, rank, world, local_rank, cores, reduce_max)
        if rank == 0:
            print(json.dumps(line))
        dist.barrier()
        dist.destroy_process_group()
        return

    sampler = ClockSampler(local_rank)
    sampler.start()
    with_cpu = main_rank and not args.no_cpu_baseline
    # per-rank barrier + max-over-ranks timing: run_step takes the barrier, the reduction happens here
    res = measure_workload_ranked(pkg, eng, wl, args, local_rank, cores, dist, reduce_max, frames, with_cpu)
    clocks = sampler.stop()

    ms_step = float(np.mean(res["times"]))
    value = world * F / (ms_step / 1000.0)
    e2e_ms = float(np.mean(res["e2e_times"]))
    e2e_value = world * F / (e2e_ms / 1000.0)
    geom, profs, deltas = res["geom"], res["profs"], res["deltas"]

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
                "cost_jobs_per_step": sum(d["cost_jobs"] for d in deltas) // max(1, len(deltas)),
                "search_launches_per_step": s_launch // max(1, len(profs)),
                "avg_launch_ms": round(s_sum_ms / max(1, s_launch), 4),
                "search_busy_ms_per_step": round(s_ms / max(1, len(profs)), 3),
                "search_us_per_job": round(1000.0 * s_ms / max(1, s_jobs), 2),
                "kernel_busy_ms_per_step": kernel_ms,
                "host_ms_per_step": {k: round(1000.0 * sum(p["host"][k] for p in profs) / len(profs), 2) for k in profs[0]["host"]},
                "note": "K4 runs out of L2 / shared memory and is bound by the integer pipe and the wavefront latency, not HBM "
                        "(SURVEY 8d); the HBM fraction is the conservative checkable figure, profiles/ holds the pipe utilisation"}

    line = {"metric": "lookahead_frames_per_s", "value": round(value, 2), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3), "ms_steps": [round(t, 1) for t in res["times"]],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u16" if depth > 8 else "u8", "data": "synthetic",
            "config": {"workload": wl["text"], "frames_per_step": F, "resolution": "%dx%d" % (W, H), "bit_depth": depth,
                       "lookahead_slices": 0, "pool_workers": cores, "async_depth": args.async_depth, "speculate": args.speculate,
                       "sm_partition": "%d SMs for cuTree / recalc / mirrors, %d for the search and cost batches (CUDA green contexts)"
                                       % tuple(geom["sm_partition"]) if geom.get("sm_partition", [0, 0])[0] else "none",
                       "streams": world, "parallelism": "independent stream per GPU" if world > 1 else "1 GPU",
                       "l2_policy": "inputs larger than L2: %d MB of pictures per step vs 126 MB L2" % (F * res["bytes_in"] // (1 << 20))},
            "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "ms_per_step": round(e2e_ms, 3), "skipped": bool(args.no_e2e),
                    "ms_steps": [round(t, 1) for t in res["e2e_times"]],
                    "h2d_bytes_per_step": int(res["e2e_delta"]["h2d"]), "d2h_bytes_per_step": int(res["e2e_delta"]["d2h"]),
                    "d2h": "per decided frame, one asynchronous mirror request into page-locked buffers: getEstimatedPictureCost, qpAqOffset, "
                           "qpCuTreeOffset, invQscaleFactor, intraCost, the coded estimate's lowresCosts + rowSatds, EVERY published "
                           "lowresMvs list (search.cpp:1968-1988), the 4 lowres planes of every non-B frame (weightPrediction.cpp:"
                           "354-365)",
                    "host_memory": "pinned" if res["pinned"] else "pageable (page-locking was refused)",
                    "host_ms_last_step": {k: round(1000.0 * v, 2) for k, v in res["e2e_prof"]["host"].items()}},
            "gpu_launches": int(sum(d["launches"] for d in deltas)),
            "clocks": clocks, "roofline": roofline,
            "decided_types": "".join(pkg.TYPE_NAMES[t][0] if t != 4 else "b" for t in res["types"][:48])}
    if "parity" in res:
        line["parity_checked"] = res["parity"]
    if "cpu_baseline" in res:
        line["cpu_baseline"] = res["cpu_baseline"]

    # --- the other BASELINE configurations, shorter runs (never allowed to break the main line) -----------------
    others = {}
    if not args.no_others and args.workload == "2160p-main10" and not args.frames:
        budget = 270.0      # seconds of wall clock the whole default run may take

        def left():
            return budget - (time.perf_counter() - t_start)
        try:
            if left() > 25:
                others["ladder-2160p-1080p-720p"] = ladder_workload(pkg, eng, args, local_rank, cores, frames[:96], rank, world, dist, reduce_max)
        except Exception as e:      # pragma: no cover
            others["ladder-2160p-1080p-720p"] = {"error": repr(e)}
        del frames
        if world > 1 and left() > 90:
            # BASELINE configs[3]: ONE 4320p stream over all the GPUs of the job (strong scaling; the 1-GPU figure is the
            # "4320p-8bit" entry of the N = 1 run)
            try:
                w4 = dict(WORKLOADS["4320p-8bit"])
                f4 = gen_frames(w4, w4["frames"], 0)
                others["4320p-window-shard"] = window_shard_line(pkg, eng, shard, dist, w4, f4, args, rank, world, local_rank, cores,
                                                                 reduce_max, steps=3, warmup=2)
                del f4
            except Exception as e:      # pragma: no cover
                others["4320p-window-shard"] = {"error": repr(e)}
        if world == 1:
            for name in ("1080p-8bit", "1080p-slower-weightp", "4320p-8bit", "1080p-8bit-hme"):
                need = 75 if name == "4320p-8bit" else 30
                if left() < need:
                    others[name] = {"skipped": "time budget of the default run (%.0f s left, %d s needed)" % (left(), need)}
                    continue
                try:
                    w2 = dict(WORKLOADS[name])
                    r2 = measure_workload(pkg, eng, w2, args, local_rank, cores, 3, 3, with_e2e=False,
                                          with_parity=with_cpu, with_cpu=with_cpu)
                    others[name] = secondary_line(w2, r2, w2["frames"])
                except Exception as e:      # pragma: no cover
                    others[name] = {"error": repr(e)}
    if others:
        line["other_workloads"] = others
    if rank == 0:
        if line.get("parity_checked", {}).get("mismatches"):
            print(json.dumps(line))
            raise SystemExit("bench.py: the CUDA lookahead differs from the reference on the timed configuration: %s"
                             % line["parity_checked"]["first_mismatches"])
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def measure_workload_ranked(pkg, eng, wl, args, device, cores, dist, reduce_max, frames, with_cpu):
    """measure_workload for the main line: every rank takes part, each step bracketed by a barrier, times reduced by max"""
    import torch
    F = wl["frames"]
    pinned_ok = [True]

    def pin(t):
        try:
            return t.pin_memory()
        except RuntimeError:
            pinned_ok[0] = False
            return t.clone()
    la_kw = dict(wl["la"], asyncDepth=args.async_depth, speculate=args.speculate, pendingMax=args.pending_max or max(8, args.async_depth),
                 batchMin=args.batch_min, device=device, poolWorkers=cores)
    host = [tuple(pin(to_t(a)) for a in f) for f in frames]
    dev = [tuple(t.cuda() for t in f) for f in host]
    torch.cuda.synchronize()
    res = {"la_kw": la_kw, "bytes_in": sum(t.numel() * t.element_size() for t in host[0])}
    for _ in range(args.warmup):
        run_step(eng, [Stream(pkg, wl, dev, la_kw, False)], dist=dist)
    times, profs, deltas, types0, geom = [], [], [], None, None
    for _ in range(args.steps):
        ms, wall, out = run_step(eng, [Stream(pkg, wl, dev, la_kw, False)], dist=dist)
        types, delta, prof, geom = out[0]
        assert types0 is None or types == types0
        times.append(reduce_max(ms)); profs.append(prof); deltas.append(delta); types0 = types
    res.update(times=times, profs=profs, deltas=deltas, types=types0, geom=geom)
    del dev
    torch.cuda.empty_cache()
    e2e_kw = dict(la_kw, extraSlots=12)
    # the end-to-end path has its own cold state (page-locked mirror buffers, the pinned-memory allocator's pools, the engine's
    # host-pointer -> device-address cache): it gets the same W untimed warm-up steps as the resident path.  Measured on B200: the
    # first two steps after ONE warm-up step still ran 514 / 414 ms against 395 ms in steady state (gpurun_out/r02w)
    for _ in range(1 if args.no_e2e else max(1, args.warmup)):
        run_step(eng, [Stream(pkg, wl, host, e2e_kw, True)], dist=dist)
    e_times = []
    for _ in range(0 if args.no_e2e else args.steps):
        ms, wall, out = run_step(eng, [Stream(pkg, wl, host, e2e_kw, True)], dist=dist)
        types, delta, prof, _ = out[0]
        assert types == types0, "decisions differ between device-resident and host-fed runs"
        e_times.append(reduce_max(ms))
    if args.no_e2e:
        e_times = list(res["times"])
    res.update(e2e_times=e_times, e2e_delta=delta, e2e_prof=prof, pinned=pinned_ok[0])
    del host
    if with_cpu:
        import refbind
        nfr = min(wl["cpu_frames"], F)
        sample = frames[:nfr]
        ok = refbind.available(wl["depth"]) and refbind.simd_selftest(wl["depth"], 6000) == 0
        r = run_reference_sample(wl, sample, cores, simd=ok, snap=True)
        if r is not None:
            fps, dt, ref, active = r
            res["parity"] = parity_check(pkg, wl, sample, ref, dict(la_kw, asyncDepth=min(16, args.async_depth)))
            ref.close()
            rc = run_reference_sample(wl, sample, cores, simd=False) if active else None
            res["cpu_baseline"] = {"value": round(fps, 3), "unit": "frames/s", "cores": cores,
                                   "kind": "reference+intrinsics" if active else "reference",
                                   "sample": "first %d frames of the same sequence, %.1f s; reference lookahead, thread pool over all %d "
                                             "host cores; %s" % (nfr, dt, cores, primitives_text(active))}
            if rc is not None:
                res["cpu_baseline"]["c_primitives_value"] = round(rc[0], 3)
                res["cpu_baseline"]["c_primitives_note"] = "the same sample with the reference's plain C primitives (its no-asm build), %.1f s" % rc[1]
        else:
            res["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "reference",
                                   "sample": "oracle/_ref not present in this snapshot"}
    return res


def window_shard_line(pkg, eng, shard, dist, wl, frames, args, rank, world, local_rank, cores, reduce_max, steps=None, warmup=None):
    """ONE stream over the GPUs of the job (SURVEY 8e level 2, strong scaling): --shard window"""
    import torch
    F = wl["frames"]
    steps = args.steps if steps is None else steps
    warmup = args.warmup if warmup is None else warmup
    la_kw = dict(wl["la"], asyncDepth=args.async_depth, speculate=args.speculate, pendingMax=args.pending_max or max(8, args.async_depth),
                 batchMin=args.batch_min, device=local_rank, poolWorkers=cores, shardCount=world)
    exchange = shard.make_exchange(dist, pkg.EXCHANGE_FN, cuda=True)
    dev = [tuple(to_t(a).cuda() for a in f) for f in frames]
    torch.cuda.synchronize()
    sh = (rank, world, exchange)
    for _ in range(warmup):
        run_step(eng, [Stream(pkg, wl, dev, la_kw, False)], dist=dist, shard=sh)
    times, types0 = [], None
    for _ in range(steps):
        ms, wall, out = run_step(eng, [Stream(pkg, wl, dev, la_kw, False)], dist=dist, shard=sh)
        times.append(reduce_max(ms)); types0 = out[0][0]
    ms = float(np.mean(times))
    return {"metric": "lookahead_frames_per_s", "value": round(F / (ms / 1000.0), 2), "unit": "frames/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u16" if wl["depth"] > 8 else "u8", "data": "synthetic",
            "config": {"workload": wl["text"], "frames_per_step": F, "pool_workers": cores,
                       "parallelism": "one stream, searches / estimates split by source frame over %d GPUs" % world},
            "decided_types": "".join(pkg.TYPE_NAMES[t][0] if t != 4 else "b" for t in types0[:48])}


if __name__ == "__main__":
    main()
