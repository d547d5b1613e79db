"""Shared parity cases: (name, depth, width, height, frames, synth kwargs, lookahead kwargs in the
reference harness' vocabulary)."""

CASES = [
    ("base8", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10)),
    ("base10", 10, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10)),
    ("pool4", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, poolThreads=4)),
    ("pool16", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, poolThreads=16)),
    ("pool16_10bit", 10, 320, 192, 60, dict(cuts=(30,)), dict(bframes=4, lookaheadDepth=20, poolThreads=16)),
    ("nob", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=0, lookaheadDepth=10)),
    ("b8", 8, 320, 192, 60, dict(cuts=(33,)), dict(bframes=8, lookaheadDepth=25)),
    ("badapt1", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=4, lookaheadDepth=12, bFrameAdaptive=1)),
    ("badapt0", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=4, lookaheadDepth=12, bFrameAdaptive=0)),
    ("nocutree", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, cuTree=0)),
    ("plain", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, cuTree=0, aqMode=0, weightp=0)),
    # aq-mode 0 without cuTree but with weightp: the qp arrays exist (weightp needs the pixel sums) and nothing ever writes them, so
    # invQscaleFactor stays 0 and every AQ-scaled cost (costEstAq, rowSatds) is 0 in the reference (slicetype.cpp:487-511, 4226-4229)
    ("aq0_weightp", 8, 320, 192, 40, dict(cuts=(20,), fades=[(5, 8, 0.3)]), dict(bframes=3, lookaheadDepth=10, cuTree=0, aqMode=0, weightp=1)),
    ("aq0_strength0_weightb10", 10, 328, 184, 30, dict(cuts=(14,)), dict(bframes=4, lookaheadDepth=12, cuTree=0, aqMode=2, aqStrength=0.0, weightp=0, weightb=1, qgSize=16)),
    # rc-lookahead barely above bframes with weightp: a frame is speculated against references the decisions have already passed and
    # released before the assumed weights are settled (found by tools/fuzz_host_vs_reference.py)
    ("shallow_la_weightp", 8, 256, 144, 36, dict(cuts=(), fades=[(6, 10, 0.3)]), dict(bframes=4, lookaheadDepth=5, weightp=1, bFrameBias=30)),
    ("shallow_la_b7_nopyramid10", 10, 256, 144, 40, dict(cuts=(19,)), dict(bframes=7, lookaheadDepth=9, bBPyramid=0, weightp=1, weightb=1, keyframeMax=25, poolThreads=2)),
    ("aq1", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, aqMode=1)),
    ("aq3", 10, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, aqMode=3)),
    ("closedgop", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, bBPyramid=0, bOpenGOP=0, keyframeMax=24, keyframeMin=2)),
    ("fade8", 8, 320, 192, 60, dict(cuts=(), fades=[(20, 12, 0.3), (40, 10, 1.0)]), dict(bframes=4, lookaheadDepth=12)),
    ("fade10", 10, 320, 192, 60, dict(cuts=(), fades=[(20, 12, 0.3), (40, 10, 1.0)]), dict(bframes=4, lookaheadDepth=12)),
    ("static", 8, 320, 192, 40, dict(cuts=(), static=True, noise=0), dict(bframes=4, lookaheadDepth=12)),
    ("static_noise", 8, 320, 192, 40, dict(cuts=(), static=True, noise=1), dict(bframes=4, lookaheadDepth=12)),
    ("static_pool", 8, 320, 192, 40, dict(cuts=(), static=True, noise=1), dict(bframes=4, lookaheadDepth=12, poolThreads=16)),
    ("flash", 8, 320, 192, 50, dict(cuts=(25,), flashes=[(10, 1), (35, 2)]), dict(bframes=4, lookaheadDepth=12)),
    ("weightb", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, weightb=1)),
    ("ragged", 8, 328, 184, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10)),
    ("small", 8, 176, 144, 30, dict(cuts=(15,)), dict(bframes=3, lookaheadDepth=10)),
    ("vbv", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, vbvBufferSize=2000, vbvMaxBitrate=2000, bitrate=1500)),
    ("sd", 8, 640, 360, 50, dict(cuts=(25,)), dict(bframes=4, lookaheadDepth=20)),
    # more of the parameter space the decisions depend on
    ("intrarefresh", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, bIntraRefresh=1, bOpenGOP=0, keyframeMax=16)),
    ("noscenecut", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=4, lookaheadDepth=12, scenecutThreshold=0)),
    ("bbias", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=4, lookaheadDepth=12, bFrameBias=40)),
    ("qg16", 10, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, qgSize=16, aqStrength=1.6)),
    ("qg64_aq1", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, qgSize=64, aqMode=1, aqStrength=0.6)),
    ("deep_la", 8, 320, 192, 70, dict(cuts=(31,)), dict(bframes=5, lookaheadDepth=40)),
    ("b16_nopyramid", 8, 320, 192, 60, dict(cuts=(), static=True, noise=2), dict(bframes=16, lookaheadDepth=30, bBPyramid=0)),
    ("fade_weightb_pool", 10, 320, 192, 60, dict(cuts=(), fades=[(15, 14, 0.25)]), dict(bframes=4, lookaheadDepth=12, weightb=1, poolThreads=16)),
    ("hd_ragged_slices", 8, 1368, 768, 20, dict(cuts=(9,)), dict(bframes=3, lookaheadDepth=10, poolThreads=8, lookaheadSlices=8)),
    ("keymin", 8, 320, 192, 50, dict(cuts=(8, 15, 22)), dict(bframes=3, lookaheadDepth=10, keyframeMax=30, keyframeMin=12)),
    # --fades: a fade-in of at least one second ends in a keyframe (bIsFadeEnd); frameVariance and the second acEnergyCu pass'
    # side effect on the weightp sums (slicetype.cpp:697-712, 1861-1906, 1972)
    ("fadein8", 8, 320, 192, 70, dict(cuts=(), envelope=[(0, 1.0), (6, 1.0), (12, 0.12), (16, 0.12), (32, 1.0), (44, 1.0), (47, 0.3), (62, 1.0)]),
     dict(bframes=4, lookaheadDepth=12, fades=1, fpsNum=10)),
    ("fadein10_sd", 10, 640, 360, 60, dict(cuts=(48,), envelope=[(0, 0.1), (3, 0.1), (20, 1.0), (30, 1.0), (34, 0.5), (46, 0.9)]),
     dict(bframes=3, lookaheadDepth=15, fades=1, fpsNum=8, weightb=1)),
    ("fadein_nowp_ragged", 8, 328, 184, 60, dict(cuts=(), envelope=[(0, 1.0), (5, 0.2), (8, 0.2), (24, 1.0)]),
     dict(bframes=4, lookaheadDepth=12, fades=1, fpsNum=10, weightp=0)),
    # --hist-scenecut (8-bit): scene cuts from per-segment histogram differences; hard cuts, a flash and a fade
    ("histcut8", 8, 640, 360, 70, dict(cuts=(17, 41), flashes=[(29, 1)],
                                        envelope=[(0, 1.0), (16, 1.0), (17, 0.5), (40, 0.5), (41, 0.95), (52, 0.95), (60, 0.3), (69, 0.3)]),
     dict(bframes=4, lookaheadDepth=16, histScenecut=1)),
    ("histcut8_ragged_nob", 8, 328, 184, 50, dict(cuts=(12, 30), envelope=[(0, 0.6), (11, 0.6), (12, 1.0), (29, 1.0), (30, 0.45), (49, 0.45)]),
     dict(bframes=0, lookaheadDepth=10, histScenecut=1)),
    ("histcut8_b8_pool", 8, 640, 360, 60, dict(cuts=(9, 22, 23, 44), envelope=[(0, 1.0), (8, 1.0), (9, 0.5), (21, 0.5), (22, 1.0), (22.5, 1.0), (23, 0.4),
                                                                               (43, 0.4), (44, 0.9), (59, 0.9)]),
     dict(bframes=8, lookaheadDepth=25, histScenecut=1, poolThreads=16)),
    # 4:0:0 (X265_CSP_I400): no chroma in the AQ energies, the weightp sums or the uploads
    ("mono400", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, csp400=1)),
    ("mono400_10_fade", 10, 320, 192, 50, dict(cuts=(), fades=[(15, 12, 0.3)]), dict(bframes=4, lookaheadDepth=12, csp400=1, aqMode=3)),
    # --temporal-layers 2: B-refs placed recursively over the mini-GOP, their costs pre-computed by compCostBref
    ("temporal2", 8, 320, 192, 60, dict(cuts=(31,), static=True, noise=2), dict(bframes=7, lookaheadDepth=20, temporalLayers=2)),
    ("temporal2_pool", 10, 320, 192, 50, dict(cuts=(24,)), dict(bframes=5, lookaheadDepth=16, temporalLayers=2, poolThreads=16)),
    # --temporal-layers 3 / 4 / 5: fixed random-access mini-GOPs of 4 / 8 / 16 pictures (Encoder::configure sets bframes 3 / 7 / 15 and
    # b-adapt 0), coded in hierarchical order; scene cuts and the end of the stream split a mini-GOP into smaller structures
    ("temporal3", 8, 320, 192, 45, dict(cuts=(14, 30)), dict(bframes=3, lookaheadDepth=12, bFrameAdaptive=0, temporalLayers=3)),
    ("temporal4_pool", 10, 320, 192, 50, dict(cuts=(21,)), dict(bframes=7, lookaheadDepth=20, bFrameAdaptive=0, temporalLayers=4, poolThreads=16)),
    ("temporal5_nopyramid_vbv", 8, 320, 192, 60, dict(cuts=(37,), static=True, noise=2),
     dict(bframes=15, lookaheadDepth=30, bFrameAdaptive=0, temporalLayers=5, bBPyramid=0, vbvBufferSize=2000, vbvMaxBitrate=2000, bitrate=1500)),
    ("temporal5", 8, 320, 192, 70, dict(cuts=(13, 41, 52)), dict(bframes=15, lookaheadDepth=35, bFrameAdaptive=0, temporalLayers=5)),
    # --radl: leading B pictures in front of the scene-cut IDRs of a closed GOP
    ("radl2", 8, 320, 192, 50, dict(cuts=(14, 31)), dict(bframes=3, lookaheadDepth=12, bOpenGOP=0, radl=2, keyframeMax=60, keyframeMin=4)),
    # slice types forced by the application (IDR, P, B runs, I) in the middle of automatic decisions
    ("forced_types", 8, 320, 192, 44, dict(cuts=(29,)), dict(bframes=3, lookaheadDepth=10)),
    # --gop-lookahead: the keyframe due at frame 20 waits for the scene cut at 22 / has nothing to wait for
    ("goplookahead_cut", 8, 320, 192, 50, dict(cuts=(22,)), dict(bframes=3, lookaheadDepth=16, keyframeMax=20, keyframeMin=2, gopLookahead=6, bOpenGOP=0)),
    ("goplookahead_nocut", 8, 320, 192, 50, dict(cuts=(33,)), dict(bframes=3, lookaheadDepth=16, keyframeMax=20, keyframeMin=2, gopLookahead=4)),
    # cooperative search slices (--lookahead-slices with a pool, >= 720 lines, no search batches)
    ("slices_badapt0", 8, 1280, 720, 24, dict(cuts=(12,)), dict(bframes=3, lookaheadDepth=10, bFrameAdaptive=0, poolThreads=4, lookaheadSlices=4)),
    ("slices_badapt1", 10, 1280, 720, 24, dict(cuts=(12,)), dict(bframes=3, lookaheadDepth=10, bFrameAdaptive=1, poolThreads=16, lookaheadSlices=3)),
    # ... next to the pool's search batches (b-adapt 2): a search is sliced or not depending on who touched it first;
    # pool 16 keeps both batches, pool 4 drops the frame-cost batch after the first decision, pool 2 the search batch too
    ("slices_trellis16", 8, 1280, 720, 30, dict(cuts=(14,)), dict(bframes=3, lookaheadDepth=10, poolThreads=16, lookaheadSlices=4)),
    ("slices_trellis4", 8, 1280, 720, 30, dict(cuts=(14,)), dict(bframes=4, lookaheadDepth=12, poolThreads=4, lookaheadSlices=4)),
    # x265's defaults (preset medium: b-adapt 2, bframes 4, rc-lookahead 20, lookahead-slices 8) with an 8-thread pool
    ("medium_defaults", 8, 1280, 720, 40, dict(cuts=(20,)), dict(bframes=4, lookaheadDepth=20, poolThreads=8, lookaheadSlices=8)),
    # narrow picture, tall enough for slices: small enough to commit as a golden fixture
    ("slices_golden", 8, 256, 720, 16, dict(cuts=(8,)), dict(bframes=3, lookaheadDepth=8, poolThreads=8, lookaheadSlices=4)),
    ("slices_trellis2", 10, 1280, 720, 30, dict(cuts=(14,), static=True, noise=1), dict(bframes=3, lookaheadDepth=10, poolThreads=2, lookaheadSlices=3)),
    # first-pass slice types of a 2-pass encode: the argument of Lookahead::addPicture (encoder.cpp:1713, 1863)
    ("pass2_types", 8, 320, 192, 44, dict(cuts=(29,)), dict(bframes=3, lookaheadDepth=10)),
    # qg-size 8: adaptive quant on 8x8 full-res blocks (4 qp offsets per lowres block); 328 / 8 and 184 / 8 are odd, which is
    # where the reference's running block index and its 2bw x 2bh addressing part ways
    ("qg8", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, qgSize=8)),
    ("qg8_ragged10", 10, 328, 184, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, qgSize=8, aqMode=3)),
    ("qg8_aq1_vbv", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, qgSize=8, aqMode=1, vbvBufferSize=2000, vbvMaxBitrate=2000, bitrate=1500)),
    # --hme (>= 540 lines): every search first runs on the 1/16-resolution planes (level 0), twice its vector is one more predictor
    # of the lowres search (level 1); per-level search method (hex / umh / dia) and range
    ("hme_default", 8, 960, 544, 16, dict(cuts=(9,)), dict(bframes=3, lookaheadDepth=8, hme=1)),
    ("hme_hexhex10", 10, 960, 540, 14, dict(cuts=(6,)), dict(bframes=2, lookaheadDepth=6, hme=1, hmeSearch0=1, hmeSearch1=1, hmeRange0=12, hmeRange1=24)),
    ("hme_umhdia_pool_fade", 8, 1024, 576, 20, dict(cuts=(), fades=[(5, 8, 0.3)]),
     dict(bframes=4, lookaheadDepth=10, hme=1, hmeSearch0=2, hmeSearch1=0, hmeRange0=20, hmeRange1=16, poolThreads=16, weightb=1)),
    # aq-mode 4 / 5 (edge): Gaussian + gradient edge map of the full-res luma, per-block edge density and mean gradient angle
    ("aq4_edge", 8, 320, 192, 30, dict(cuts=(15,)), dict(bframes=3, lookaheadDepth=10, aqMode=4)),
    ("aq5_edge10_ragged", 10, 328, 184, 30, dict(cuts=(15,)), dict(bframes=3, lookaheadDepth=10, aqMode=5, aqStrength=1.3)),
    ("aq4_qg8_sd", 8, 640, 360, 24, dict(cuts=(11,)), dict(bframes=4, lookaheadDepth=12, aqMode=4, qgSize=8, poolThreads=16)),
    ("hme_star", 8, 960, 544, 14, dict(cuts=(7,), n_rects=8), dict(bframes=3, lookaheadDepth=8, hme=1, hmeSearch0=3, hmeSearch1=3, hmeRange0=16, hmeRange1=32)),
    ("hme_fullhex", 8, 960, 544, 10, dict(cuts=(5,)), dict(bframes=2, lookaheadDepth=6, hme=1, hmeSearch0=5, hmeSearch1=1, hmeRange0=8, hmeRange1=16)),
    ("hme_hexstar10_pool", 10, 960, 540, 12, dict(cuts=(5,)), dict(bframes=2, lookaheadDepth=6, hme=1, hmeSearch0=1, hmeSearch1=3, hmeRange1=16, poolThreads=16)),
    # 12-bit (main12): 16-bit samples whose SATD coefficients no longer fit the packed 16-bit lanes of the 8 / 10-bit kernels
    ("base12", 12, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10)),
    ("fade12_pool_weightb", 12, 320, 192, 50, dict(cuts=(), fades=[(15, 12, 0.3), (35, 9, 1.0)]), dict(bframes=4, lookaheadDepth=12, poolThreads=16, weightb=1, aqMode=3)),
    ("hme12_aq4", 12, 960, 544, 12, dict(cuts=(6,)), dict(bframes=2, lookaheadDepth=6, hme=1, aqMode=4)),
    # short enough to commit as a golden fixture
    ("hme_golden", 8, 960, 544, 8, dict(cuts=(4,)), dict(bframes=2, lookaheadDepth=5, hme=1)),
    ("vbv_nocutree", 8, 320, 192, 40, dict(cuts=(20,)), dict(bframes=3, lookaheadDepth=10, cuTree=0, vbvBufferSize=2000, vbvMaxBitrate=2000, bitrate=1500)),
]

# cases that also run Lookahead::getEstimatedPictureCost (+ its VBV row aggregation) on every decided frame, the way
# Encoder::encode does (encoder.cpp:2367), and compare satdCost / satdForVbv / the rescaled arrays
ESTIMATE = ["base8", "base10", "vbv", "radl2", "nocutree", "plain", "intrarefresh", "closedgop", "qg8", "qg8_aq1_vbv", "vbv_nocutree",
            "b8", "pool16"]
# intra-refresh column range handed to the VBV aggregation (FrameData::m_pir, slicetype.cpp:1425-1427)
PIR = {"intrarefresh": (2, 3)}

# subset small enough to commit as golden fixtures and to run in the quick CPU suite
GOLDEN = ["base8", "base10", "pool16", "fade8", "static_noise", "ragged", "nob", "slices_golden", "hme_golden", "aq4_edge", "temporal3"]

REF2LA = dict(bframes="bframes", lookaheadDepth="lookaheadDepth", bFrameAdaptive="bFrameAdaptive", bBPyramid="bBPyramid",
              scenecutThreshold="scenecutThreshold", keyframeMax="keyframeMax", keyframeMin="keyframeMin",
              bOpenGOP="bOpenGOP", aqMode="aqMode", aqStrength="aqStrength", cuTree="cuTree", qCompress="qCompress",
              weightp="bEnableWeightedPred", weightb="bEnableWeightedBiPred", qgSize="qgSize", bFrameBias="bFrameBias",
              scenecutBias="scenecutBias", vbvBufferSize="vbvBufferSize", vbvMaxBitrate="vbvMaxBitrate",
              poolThreads="poolWorkers", lookaheadSlices="lookaheadSlices", gopLookahead="gopLookahead", bIntraRefresh="bIntraRefresh", radl="radl",
              fades="bEnableFades", fpsNum="fpsNum", fpsDenom="fpsDenom", temporalLayers="bEnableTemporalSubLayers", histScenecut="bHistBasedSceneCut", hme="bEnableHME")


def la_kwargs(refkw):
    kw = {REF2LA[k]: v for k, v in refkw.items() if k in REF2LA}
    if "bitrate" in refkw:
        kw["rateControlMode"] = 0
    if refkw.get("hme"):
        kw["hmeSearchMethod"] = (refkw.get("hmeSearch0", 1), refkw.get("hmeSearch1", 2))
        kw["hmeRange"] = (refkw.get("hmeRange0", 16), refkw.get("hmeRange1", 32))
    return kw


def get_case(name):
    for c in CASES:
        if c[0] == name:
            return c
    raise KeyError(name)


def make_seq(synth, case):
    name, depth, w, h, n, skw, _ = case
    return synth.SynthSequence(w, h, depth=depth, seed=1, **skw)


# application-forced slice types (x265_picture::sliceType -> Lowres::sliceTypeReq) per case: {poc: X265_TYPE_*};
# 1 IDR, 2 I, 3 P, 4 BREF, 5 B
FORCED = {
    "forced_types": {9: 1, 16: 3, 17: 3, 23: 5, 24: 5, 25: 3, 33: 2},
}
# first-pass types handed to addPicture (2-pass): every frame typed, like a stats file would
PASS2 = {
    "pass2_types": {i: t for i, t in enumerate([1, 5, 4, 5, 3, 5, 5, 3, 3, 5, 4, 5, 3, 5, 3, 3, 5, 5, 5, 3, 5, 3, 5, 4, 5, 3, 5, 5, 3, 2, 5,
                                                5, 3, 3, 5, 4, 5, 3, 5, 3, 5, 5, 3, 3])},
}

# cases only the CPU suite runs (host logic through the sim engine against the live reference)
CPU_ONLY = []


def run_reference(refbind, synth, case, planes=True, estimate=None):
    name, depth, w, h, n, skw, rkw = case
    seq = make_seq(synth, case)
    forced, pass2 = FORCED.get(name, {}), PASS2.get(name, {})
    estimate = (name in ESTIMATE) if estimate is None else estimate
    ref = refbind.RefLookahead(w, h, depth=depth, dumpPlanes=1 if planes else 0,
                               keepFrames=(rkw.get("lookaheadDepth", 20) + 3 * (rkw.get("bframes", 4) + 2)) if estimate else 0, **rkw)
    pir = PIR.get(name, (0, 0))
    import _pkg
    tracker = _pkg.load_pkg().RefTracker()
    done = [0]

    def estimate_new():
        # getEstimatedPictureCost on the frames decided since the last call, while their references are alive
        while estimate and done[0] < ref.num_out():
            poc, t = ref.out_info(done[0])
            r0, r1 = tracker.push(poc, t, done[0])
            ref.estimate(done[0], -1 if r0 is None else r0, -1 if r1 is None else r1, pir)
            done[0] += 1
    for i in range(n):
        ref.put(*seq.frame(i), slice_type=forced.get(i, 0), pass2_type=pass2.get(i, 0))
        estimate_new()
    ref.flush()
    estimate_new()
    out = ref.frames()
    ref.close()
    return out


def run_ours(pkg, synth, case, lib_path=None, planes=True, **extra):
    name, depth, w, h, n, skw, rkw = case
    seq = make_seq(synth, case)
    kw = la_kwargs(rkw)
    kw.update(extra)
    la = pkg.Lookahead(w, h, depth=depth, lib_path=lib_path, **kw)
    mono = bool(rkw.get("csp400"))      # 4:0:0: the pictures carry no chroma planes
    out = pkg.run_sequence(la, ((seq.frame(i)[0], None, None) if mono else seq.frame(i) for i in range(n)), planes=planes, slice_types=FORCED.get(name),
                           pass2_types=PASS2.get(name), estimate_cost=name in ESTIMATE, pir=PIR.get(name, (-1, -1)))
    la.close()
    return out


# BASELINE.json's own configurations at full size (VERDICT r01 "N2"): (name, depth, w, h, frames, synth kwargs, lookahead kwargs).
# The pool size is part of the configuration (x265 runs a pool of all cores by default; >= 13 workers keeps both of the
# reference's batch modes, slicetype.cpp:2691,2733), cooperative slices are off as in every parity run.
FULL_SIZE = [
    ("cfg2_2160p_main10", 10, 3840, 2160, 72, dict(cuts=(39,), n_rects=6, seed=2), dict(bframes=8, lookaheadDepth=60, poolThreads=16)),
    ("cfg1_1080p_8bit", 8, 1920, 1080, 100, dict(cuts=(53,), n_rects=6, seed=1), dict(bframes=4, lookaheadDepth=20, poolThreads=16)),
    ("cfg3_1080p_slower_weightp", 8, 1920, 1080, 90, dict(cuts=(13, 37, 64), fades=[(20, 12, 0.35), (70, 10, 1.0)], flashes=[(50, 1)], n_rects=6, seed=3),
     dict(bframes=8, lookaheadDepth=40, poolThreads=16, weightp=1)),
    # --hme at a BASELINE size: level-0 (hex) and level-1 (umh) vectors and costs of every published search included
    # aq-mode 4 (edge map of the full-res luma) + three temporal layers (random-access mini-GOPs of 4, split at the cut) at a BASELINE size
    ("cfg1_1080p_8bit_aq4_tl3", 8, 1920, 1080, 38, dict(cuts=(18,), n_rects=6, seed=5),
     dict(bframes=3, lookaheadDepth=20, bFrameAdaptive=0, temporalLayers=3, aqMode=4, poolThreads=16)),
    ("cfg1_1080p_8bit_hme", 8, 1920, 1080, 40, dict(cuts=(21,), n_rects=6, seed=1), dict(bframes=4, lookaheadDepth=20, poolThreads=16, hme=1)),
]
# config 4 (7680x4320, rc-lookahead 80): against the C oracle through the sim engine on a dozen frames
FULL_SIZE_ORACLE = ("cfg4_4320p_8bit", 8, 7680, 4320, 12, dict(cuts=(7,), n_rects=4, seed=4), dict(bframes=4, lookaheadDepth=80))


def compare_streaming(pkg, synth, refbind, compare, case, frames=None, **extra):
    """Full-size parity without holding two copies of every array: the reference runs first (its snapshots stay inside
    the harness), then ours, each decided frame compared against the reference's and dropped.  Returns
    (mismatch strings, frames compared)."""
    name, depth, w, h, n, skw, rkw = case
    skw = dict(skw)
    seq = synth.SynthSequence(w, h, depth=depth, **skw)
    if frames is None:
        frames = [seq.frame(i) for i in range(n)]
    ref = refbind.RefLookahead(w, h, depth=depth, **rkw)
    for f in frames:
        ref.put(*f)
    ref.flush()
    kw = la_kwargs(rkw)
    kw.update(extra)
    la = pkg.Lookahead(w, h, depth=depth, **kw)
    bad, idx = [], [0]

    def drain():
        while True:
            info = la.get_decided()
            if info is None:
                return
            got = la.frame_dict(info, planes=False)
            la.release(info.handle)
            want = ref.frame(idx[0], drop=True)
            idx[0] += 1
            if len(bad) < 20:
                bad.extend(compare.compare_frames(want, got, cutree=rkw.get("cuTree", 1), weightp=rkw.get("weightp", 1)))
    for i, (y, u, v) in enumerate(frames):
        la.add_picture(y, u, v, pts=i)
        drain()
    la.flush()
    drain()
    la.close()
    nref = ref.num_out()
    ref.close()
    if idx[0] != nref or nref != len(frames):
        bad.append("frame count: reference %d, ours %d, input %d" % (nref, idx[0], len(frames)))
    return bad, idx[0]
