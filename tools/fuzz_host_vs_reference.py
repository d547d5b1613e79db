"""Randomised parity sweep of the PRODUCT's host logic (through the CPU sim engine) against the live unmodified reference:
seeded random lookahead configurations on small synthetic sequences, every published field compared (tests/compare.py).
CPU only; needs oracle/_ref.   python tools/fuzz_host_vs_reference.py [n_cases] [first_seed]   or   ... seeds 3,17,254"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import _pkg
import build_sim
import cases
import compare
import refbind


def _engine_lib(simdir, depth):
    """the CPU sim engine, or with FUZZ_ENGINE=cuda the real library on a GPU box (lib_path None = x265-amod_b200/lib)"""
    if os.environ.get("FUZZ_ENGINE") == "cuda":
        return None
    return os.path.join(simdir, "libx265la_sim%d.so" % depth)


def random_case(seed):
    r = np.random.default_rng(seed)
    depth = int(r.choice([8, 8, 10, 12]))
    w, h = [(176, 144), (320, 192), (328, 184), (256, 144)][int(r.integers(4))]
    n = int(r.integers(12, 46))
    tl = int(r.choice([0, 0, 0, 0, 2, 3, 4]))
    bframes = int(r.integers(0, 9))
    badapt = int(r.choice([0, 1, 2, 2]))
    if tl > 2:
        bframes, badapt = {3: 3, 4: 7}[tl], 0
    if tl == 2 and bframes < 2:
        tl = 0
    la = dict(bframes=bframes, bFrameAdaptive=badapt, lookaheadDepth=int(r.integers(bframes + 1, bframes + 22)),
              bBPyramid=int(r.integers(2)), bOpenGOP=int(r.integers(2)), scenecutThreshold=int(r.choice([0, 40, 40, 60])),
              keyframeMax=int(r.choice([250, 250, 12, 25, 40])), aqMode=int(r.integers(0, 6)), aqStrength=float(r.choice([0.0, 0.6, 1.0, 1.5])),
              cuTree=int(r.integers(2)), weightp=int(r.integers(2)), weightb=int(r.integers(2)), poolThreads=int(r.choice([0, 0, 2, 4, 16])),
              qgSize=int(r.choice([8, 16, 32, 64])), bFrameBias=int(r.choice([0, 0, -20, 30])), temporalLayers=tl)
    la["keyframeMin"] = int(r.choice([0, 0, 2, 8])) if la["keyframeMax"] > 8 else 0
    if r.integers(5) == 0:
        la.update(vbvBufferSize=2000, vbvMaxBitrate=2000, bitrate=1500)
    if r.integers(6) == 0 and not la["bOpenGOP"]:
        la["radl"] = int(r.integers(1, 3))
    if r.integers(6) == 0:
        la["gopLookahead"] = int(r.integers(1, 6))
    if r.integers(6) == 0 and la["aqMode"] < 4 and la["qgSize"] != 8 and (la["aqMode"] or la["weightp"] or la["weightb"]):
        la.update(fades=1, fpsNum=int(r.choice([8, 10, 25])))
    if la.get("fades") and la["lookaheadDepth"] < la["bframes"] + 2:
        # the fade detector walks the first bframes + 2 frames of the input queue at every decision (slicetype.cpp:1861-1906): with a
        # shorter rc-lookahead, how many it finds depends on WHEN a pool worker got to run slicetypeDecide -- timing, not parameters
        la["lookaheadDepth"] = la["bframes"] + 2
    if r.integers(6) == 0 and depth == 8 and w >= 256:
        la["histScenecut"] = 1
    if r.integers(8) == 0:
        la["bIntraRefresh"] = 1
    if r.integers(8) == 0:
        la["csp400"] = 1
    if seed >= 1000000:
        # "big" seeds: pictures tall enough for cooperative lookahead slices (>= 720 lines, with a pool) or --hme (>= 540 lines), short
        big = int(r.integers(3))
        n = int(r.integers(9, 17))
        la["lookaheadDepth"] = min(la["lookaheadDepth"], la["bframes"] + 8)
        for k in ("histScenecut", "fades", "fpsNum", "csp400"):
            la.pop(k, None)
        if la["temporalLayers"] > 2:
            la["temporalLayers"] = 0; la["bFrameAdaptive"] = int(r.integers(3)); la["bframes"] = min(la["bframes"], 4)
            la["lookaheadDepth"] = la["bframes"] + int(r.integers(2, 8))
        if big == 0:
            w, h = 1280, 720
            la["poolThreads"] = int(r.choice([2, 4, 8, 16])); la["lookaheadSlices"] = int(r.choice([2, 3, 4, 8]))
        elif big == 1:
            w, h = 960, 544
            la.update(hme=1, hmeSearch0=int(r.choice([0, 1, 2, 3])), hmeSearch1=int(r.choice([0, 1, 2, 2, 3])),
                      hmeRange0=int(r.choice([8, 16, 24])), hmeRange1=int(r.choice([16, 32])))
        else:
            w, h = 1024, 576
    cuts = tuple(sorted(set(int(x) for x in r.integers(3, n, int(r.integers(0, 3))))))
    skw = dict(cuts=cuts)
    kind = int(r.integers(5))
    if kind == 0:
        skw.update(static=True, noise=int(r.integers(0, 3)))
    elif kind == 1:
        a = int(r.integers(3, max(4, n - 12))); skw.update(fades=[(a, int(r.integers(5, 12)), float(r.choice([0.25, 0.4, 1.0])))])
    elif kind == 2:
        skw.update(flashes=[(int(r.integers(2, n - 2)), int(r.integers(1, 3)))])
    return ("fuzz%d" % seed, depth, w, h, n, skw, la)


def _case_setup(seed):
    case = random_case(seed)
    name, depth, w, h, n, skw, la = case
    # one case in three also drives Lookahead::getEstimatedPictureCost (+ the VBV row sums) on every decided frame, like Encoder::encode
    # (not with temporal layers: RefTracker models the nearest references, not the layered reference picture sets)
    estimate = seed % 3 == 0 and not la.get("temporalLayers", 0) and not la.get("radl") and not la.get("bIntraRefresh")
    if estimate:
        cases.ESTIMATE.append(name)
    if seed % 7 == 3 and not la.get("temporalLayers", 0):
        # slice types forced by the application (x265_picture::sliceType -> Lowres::sliceTypeReq) on a few frames
        r = np.random.default_rng(seed + 77)
        pocs = sorted(set(int(x) for x in r.integers(2, n - 1, int(r.integers(1, 5)))))
        cases.FORCED[name] = {p: int(r.choice([1, 2, 3, 3, 5, 5])) for p in pocs}
    return case, estimate


def run_reference_side(synth, seed, path):
    """child 1: the reference's run, pickled to `path`.  0 = done, 2 = it refused the configuration (a crash shows as a signal)"""
    import pickle
    case, estimate = _case_setup(seed)
    if not refbind.available(case[1]):
        return 2
    pass2 = None
    try:
        if seed % 11 == 5 and not case[6].get("temporalLayers", 0) and case[0] not in cases.FORCED:
            # a 2-pass encode: the types the reference decided in a first run come back as the argument of Lookahead::addPicture
            first = cases.run_reference(refbind, synth, case, estimate=False, planes=False)
            pass2 = {f["poc"]: f["sliceType"] for f in first}
            cases.PASS2[case[0]] = pass2
        want = cases.run_reference(refbind, synth, case, estimate=estimate)
    except Exception as e:
        print(case[0], "reference refused:", repr(e)[:100]); return 2
    with open(path, "wb") as f:
        pickle.dump((want, pass2), f)
    return 0


def run_our_side(pkg, synth, simdir, seed, path):
    """child 2: the product's run through the sim engine, compared with the pickled reference run.
    0 = identical, 1 = mismatch, 2 = the host library refused the configuration"""
    import pickle
    case, estimate = _case_setup(seed)
    name, depth, w, h, n, skw, la = case
    with open(path, "rb") as f:
        want, pass2 = pickle.load(f)
    if pass2:
        cases.PASS2[name] = pass2
    try:
        # how the product schedules its GPU work must not show in the results: random scheduling mode and extra input delay
        r = np.random.default_rng(seed + 1000003)
        extra = dict(speculate=int(r.choice([0, 1, 1, 2])), asyncDepth=int(r.choice([0, 0, 3, 9, 17])))
        got = cases.run_ours(pkg, synth, case, lib_path=_engine_lib(simdir, depth), **extra)
    except RuntimeError as e:
        print(name, "REFUSED by the host library:", str(e)[:160], la); return 2
    bad = compare.compare_runs(want, got, check_planes=True, cutree=la.get("cuTree", 1), weightp=la.get("weightp", 1) or la.get("weightb", 0),
                               vbv=bool(la.get("vbvBufferSize")), skip_propagate=tuple(cases.FORCED.get(name, {})) + tuple(pass2 or ()))
    # a B frame of the analysis that slicetypeDecide turns into the P in front of an IDR (closed GOP, slicetype.cpp:2012-2016) was
    # a B frame to cuTree: its propagateCost is whatever the allocation held, like that of a forced-type frame (compare.py)
    idr_pocs = set(f["poc"] for f in want if f["sliceType"] == 1)
    bad = [b for b in bad if not ("propagateCost" in b and any(("poc=%d:" % (p - 1)) in b for p in idr_pocs))]
    # ... and the reference leaves propagateCost of a keyframe unwritten when its re-analysis finds too few frames behind it (end of
    # the stream): the values differ from run to run of the reference itself.  Reported, not counted
    soft = [b for b in bad if "propagateCost" in b]
    bad = [b for b in bad if "propagateCost" not in b]
    if soft and not bad:
        print(name, "(propagateCost only, undefined in the reference there):", soft[0].strip())
    if bad:
        print(name, "MISMATCH", depth, w, h, n, skw, la)
        print("   ", "\n    ".join(bad[:4]))
        return 1
    return 0


def _child(fn, *args):
    sys.stdout.flush()
    pid = os.fork()
    if pid == 0:
        rc = 3
        try:
            rc = fn(*args)
        finally:
            sys.stdout.flush()
            os._exit(rc)
    _, status = os.waitpid(pid, 0)
    return os.WEXITSTATUS(status) if os.WIFEXITED(status) else -1


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "seeds":
        seeds = [int(x) for x in sys.argv[2].split(",")]
    else:
        n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 50
        first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        seeds = list(range(first, first + n_cases))
    n_cases = len(seeds)
    pkg = _pkg.load_pkg(); synth = _pkg.load_synth()
    simdir = build_sim.build()
    path = "/tmp/fuzz_ref_%d.pkl" % os.getpid()
    c = dict(identical=0, mismatch=0, ref_refused=0, ref_crashed=0, ours_refused=0, ours_crashed=0)
    t0 = time.time()
    for seed in seeds:
        # one child per side and case: the reference itself crashes on a few combinations (hist-scenecut on 4:0:0, --radl with a scene
        # cut in the last frames, ...), and a crash of the PRODUCT must be told apart from that
        rc = _child(run_reference_side, synth, seed, path)
        if rc != 0:
            c["ref_refused" if rc == 2 else "ref_crashed"] += 1
            if rc != 2:
                # the product must still survive (or refuse) what kills the reference
                rc2 = _child(lambda: 0 if _survives(pkg, synth, simdir, seed) else 0)
                if rc2 != 0:
                    c["ours_crashed"] += 1
                    print("fuzz%d: THE PRODUCT DIED where the reference dies too: %s" % (seed, random_case(seed)[1:]))
            continue
        rc = _child(run_our_side, pkg, synth, simdir, seed, path)
        if rc == 0: c["identical"] += 1
        elif rc == 1: c["mismatch"] += 1
        elif rc == 2: c["ours_refused"] += 1
        else:
            c["ours_crashed"] += 1
            print("fuzz%d: THE PRODUCT DIED: %s" % (seed, random_case(seed)[1:]))
    if os.path.exists(path):
        os.remove(path)
    print("%d cases: %s, %.0f s" % (n_cases, ", ".join("%s %d" % kv for kv in c.items()), time.time() - t0))
    return 1 if c["mismatch"] or c["ours_crashed"] else 0


def _survives(pkg, synth, simdir, seed):
    case, _ = _case_setup(seed)
    try:
        cases.run_ours(pkg, synth, case, lib_path=_engine_lib(simdir, case[1]))
    except RuntimeError:
        pass        # a refusal is fine
    return True


if __name__ == "__main__":
    sys.exit(main())
