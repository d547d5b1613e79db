"""Field-by-field comparison of decided-frame dicts (oracle/refbind.py layout)."""
import numpy as np

EXACT_SCALARS = ["poc", "sliceType", "bScenecut", "bKeyframe", "bLastMiniGopBFrame", "leadingBframes"]


def compare_frames(ref, got, check_planes=False, qp_tol=1e-3, weightp=True, cutree=True, label="", skip_propagate=(), vbv=False,
                   decision_rank=True):
    """decision_rank=False: `got` comes from a rank > 0 of a sharded stream, which takes the same decisions but does not run
    cuTree (its qp offsets are never read): qpCuTreeOffset / propagateCost are not compared"""
    """ref: dict from the reference harness; got: dict from our Lookahead.  Returns list of
    mismatch strings (empty = parity)."""
    bad = []
    tag = "%sframe poc=%d: " % (label, ref["poc"])
    for k in EXACT_SCALARS:
        if int(ref[k]) != int(got[k]):
            bad.append(tag + "%s ref=%s got=%s" % (k, ref[k], got[k]))
    if bad:
        return bad
    nb = ref["nb"]
    if not np.array_equal(ref["costEst"], got["costEst"]):
        bad.append(tag + "costEst\nref=%s\ngot=%s" % (ref["costEst"], got["costEst"]))
    if not np.array_equal(ref["costEstAq"], got["costEstAq"]):
        bad.append(tag + "costEstAq\nref=%s\ngot=%s" % (ref["costEstAq"], got["costEstAq"]))
    if not np.array_equal(ref["intraMbs"], got["intraMbs"]):
        bad.append(tag + "intraMbs ref=%s got=%s" % (ref["intraMbs"], got["intraMbs"]))
    for k in ("intraCost", "intraMode", "invQscaleFactor"):
        if not np.array_equal(ref[k], got[k]):
            n = int(np.sum(ref[k] != got[k]))
            bad.append(tag + "%s differs in %d blocks" % (k, n))
    for k in ("qpAqOffset", "qpCuTreeOffset"):
        if k == "qpCuTreeOffset" and not decision_rank:
            continue
        d = np.max(np.abs(ref[k] - got[k])) if len(ref[k]) else 0.0
        if not d <= qp_tol:
            bad.append(tag + "%s max|delta|=%g" % (k, d))
    # propagateCost is only defined for I/P frames: the reference never initialises it for B frames, and
    # for B-refs cuTree resets frames[curnonb + (bframes+1)/2] (slicetype.cpp:3460) while placeBref marks
    # list[bframes/2] (:1757), which are different frames for even mini-GOP sizes.
    # ... and a frame whose type the application forced (Lowres::sliceTypeReq) is analysed as AUTO: when the analysis made it
    # a B frame and slicetypeDecide then imposes P on it (slicetype.cpp:1938), cuTree only ever cleared its first row and
    # the rest is whatever the malloc'ed (lowres.cpp: CHECKED_MALLOC, not zeroed) array held.
    # ... and with --temporal-layers 3..5 the P that closes a sub-structure of a SPLIT mini-GOP (slicetype.cpp:2084-2090) was a B
    # frame to the analysis and to cuTree, too
    split_p = "gopIdTop" in got and got["gopId"] != got["gopIdTop"]
    if cutree and decision_rank and ref["sliceType"] in (1, 2, 3) and ref["poc"] not in skip_propagate and not split_p and \
            not np.array_equal(ref["propagateCost"], got["propagateCost"]):
        n = int(np.sum(ref["propagateCost"] != got["propagateCost"]))
        bad.append(tag + "propagateCost differs in %d blocks" % n)
    if "tempLayer" in got and "tempLayer" in ref:       # --temporal-layers 3..5: what the DPB reads of the decision
        for k in ("tempLayer", "gopOffset") + (("gopId",) if got["gopId"] is not None else ()):
            if int(ref[k]) != int(got[k]):
                bad.append(tag + "%s ref=%d got=%d" % (k, ref[k], got[k]))
    if "bIsFadeEnd" in ref and "bIsFadeEnd" in got:
        if int(ref["bIsFadeEnd"]) != int(got["bIsFadeEnd"]):
            bad.append(tag + "bIsFadeEnd ref=%d got=%d" % (ref["bIsFadeEnd"], got["bIsFadeEnd"]))
        if ref.get("frameVariance", 0) != got.get("frameVariance", 0):
            bad.append(tag + "frameVariance ref=%r got=%r" % (ref.get("frameVariance"), got.get("frameVariance")))
    if "histCheck" in got and ref.get("histCheck"):
        for k in ("histVar", "histAvg", "histCheck"):
            if ref[k] != got[k]:
                bad.append(tag + "%s ref=%s got=%s" % (k, ref[k], got[k]))
    if weightp:
        if not np.array_equal(ref["wp_ssd"], got["wp_ssd"]) or not np.array_equal(ref["wp_sum"], got["wp_sum"]):
            bad.append(tag + "wp stats ref=%s/%s got=%s/%s" % (ref["wp_ssd"], ref["wp_sum"], got["wp_ssd"], got["wp_sum"]))
    for l in range(2):
        for d in range(nb):
            ref_searched = ref["mvs"][l, d, 0, 0] != 0x7FFF
            if bool(ref_searched) != bool(got["searched"][l, d]):
                bad.append(tag + "mv sentinel list=%d dist=%d ref_searched=%s got=%s" % (l, d, ref_searched, got["searched"][l, d]))
                continue
            if ref_searched:
                if not np.array_equal(ref["mvs"][l, d], got["mvs"][l, d]):
                    n = int(np.sum(np.any(ref["mvs"][l, d] != got["mvs"][l, d], axis=1)))
                    bad.append(tag + "lowresMvs[%d][%d] differ in %d blocks" % (l, d, n))
                if not np.array_equal(ref["mvCosts"][l, d], got["mvCosts"][l, d]):
                    n = int(np.sum(ref["mvCosts"][l, d] != got["mvCosts"][l, d]))
                    bad.append(tag + "lowresMvCosts[%d][%d] differ in %d blocks" % (l, d, n))
                if "lowerMvs" in ref and "lowerMvs" in got:     # --hme: the level-0 search behind it
                    if not np.array_equal(ref["lowerMvs"][l, d], got["lowerMvs"][l, d]):
                        n = int(np.sum(np.any(ref["lowerMvs"][l, d] != got["lowerMvs"][l, d], axis=1)))
                        bad.append(tag + "lowerResMvs[%d][%d] differ in %d blocks" % (l, d, n))
                    if not np.array_equal(ref["lowerMvCosts"][l, d], got["lowerMvCosts"][l, d]):
                        n = int(np.sum(ref["lowerMvCosts"][l, d] != got["lowerMvCosts"][l, d]))
                        bad.append(tag + "lowerResMvCosts[%d][%d] differ in %d blocks" % (l, d, n))
    for i in range(nb):
        for j in range(nb):
            ref_done = ref["rowSatds"][i, j, 0] != -1 and ref["costEst"][i, j] >= 0
            if not ref_done:
                continue
            if not np.array_equal(ref["lowresCosts"][i, j], got["lowresCosts"][i, j]):
                n = int(np.sum(ref["lowresCosts"][i, j] != got["lowresCosts"][i, j]))
                bad.append(tag + "lowresCosts[%d][%d] differ in %d blocks" % (i, j, n))
            # non-B frames' rowSatds of the coded estimate are rewritten by frameCostRecalculate
            if not np.array_equal(ref["rowSatds"][i, j], got["rowSatds"][i, j]):
                bad.append(tag + "rowSatds[%d][%d] differ" % (i, j))
    # what RateControl reads of the VBV lookahead (ratecontrol.cpp:2512-2577)
    if vbv and "plannedType" in ref and "plannedType" in got:
        if int(ref["indB"]) != int(got["indB"]):
            bad.append(tag + "indB ref=%d got=%d" % (ref["indB"], got["indB"]))
        nt = int(np.argmax(ref["plannedType"] == 0)) if np.any(ref["plannedType"] == 0) else len(ref["plannedType"])
        nt = max(nt, int(ref["indB"]))
        if not np.array_equal(ref["plannedType"][:nt], got["plannedType"][:nt]):
            bad.append(tag + "plannedType ref=%s got=%s" % (ref["plannedType"][:nt], got["plannedType"][:nt]))
        if not np.array_equal(ref["plannedSatd"][:nt], got["plannedSatd"][:nt]):
            bad.append(tag + "plannedSatd ref=%s got=%s" % (ref["plannedSatd"][:nt], got["plannedSatd"][:nt]))
    # Lookahead::getEstimatedPictureCost as the encoder calls it per coded frame
    if "est" in ref and "est" in got:
        e, g = ref["est"], got["est"]
        if int(e["satdCost"]) != int(g["satdCost"]):
            bad.append(tag + "satdCost ref=%d got=%d" % (e["satdCost"], g["satdCost"]))
        if "rowSatds" in g and not np.array_equal(e["rowSatds"], g["rowSatds"]):
            bad.append(tag + "rowSatds after getEstimatedPictureCost differ")
        if vbv and "satdForVbv" in g:
            for k in ("satdForVbv", "intraSatdForVbv", "lowresCostForRc", "intraCostForRc"):
                if not np.array_equal(e[k], g[k]):
                    bad.append(tag + "%s differs in %d entries" % (k, int(np.sum(e[k] != g[k]))))
    elif "est" in ref:
        bad.append(tag + "getEstimatedPictureCost ran on the reference side only")
    if check_planes and "planes" in ref and "planes" in got:
        if not np.array_equal(ref["planes"], got["planes"]):
            n = int(np.sum(ref["planes"] != got["planes"]))
            bad.append(tag + "lowres planes differ in %d samples" % n)
    return bad


def compare_runs(ref_frames, got_frames, **kw):
    bad = []
    if len(ref_frames) != len(got_frames):
        bad.append("frame count ref=%d got=%d" % (len(ref_frames), len(got_frames)))
    for r, g in zip(ref_frames, got_frames):
        bad += compare_frames(r, g, **kw)
        if len(bad) > 20:
            break
    return bad
