"""Golden fixtures: decided-frame dumps of the UNMODIFIED reference (oracle/_ref), stored as
compressed .npz under tests/golden/ by tests/golden/make_golden.py."""
import os
import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KEYS = ["costEst", "costEstAq", "intraMbs", "rowSatds", "lowresCosts", "mvs", "mvCosts", "intraCost", "intraMode",
        "qpAqOffset", "qpCuTreeOffset", "invQscaleFactor", "propagateCost", "wp_ssd", "wp_sum", "weightedCostDelta"]
SCALARS = ["poc", "sliceType", "bScenecut", "bKeyframe", "bLastMiniGopBFrame", "leadingBframes", "bw", "bh", "nb",
           "stride", "planeLines"]


def path(name):
    return os.path.join(HERE, name + ".npz")


def save(name, frames):
    d = {"n": np.array(len(frames))}
    for i, f in enumerate(frames):
        d["s%d" % i] = np.array([int(f[k]) for k in SCALARS], np.int64)
        f = dict(f)
        # entries the reference never computed hold uninitialised memory: blank them so the
        # fixtures are deterministic and compress well (the comparison skips them anyway)
        lc = f["lowresCosts"].copy(); rs = f["rowSatds"].copy(); mv = f["mvs"].copy(); mc = f["mvCosts"].copy()
        nb = f["nb"]
        for a in range(nb):
            for b in range(nb):
                if not (rs[a, b, 0] != -1 and f["costEst"][a, b] >= 0):
                    lc[a, b] = 0
                    keep = rs[a, b, 0]
                    rs[a, b] = 0
                    rs[a, b, 0] = keep
        for l in range(2):
            for a in range(nb):
                if mv[l, a, 0, 0] == 0x7FFF:
                    mv[l, a] = 0
                    mv[l, a, 0, 0] = 0x7FFF
                    mc[l, a] = 0
        if f["sliceType"] not in (1, 2, 3):
            f["propagateCost"] = np.zeros_like(f["propagateCost"])
        f["lowresCosts"], f["rowSatds"], f["mvs"], f["mvCosts"] = lc, rs, mv, mc
        for k in KEYS:
            a = f[k]
            if k == "mvs":
                a = a.astype(np.int16)       # lowres MVs and the 0x7FFF sentinel fit int16
            d["%s%d" % (k, i)] = a
        # planes are large: keep a checksum only
        if "planes" in f:
            d["planesum%d" % i] = np.array([int(np.sum(f["planes"].astype(np.uint64) * (np.arange(f["planes"].size, dtype=np.uint64).reshape(f["planes"].shape) % 251 + 1)))], np.uint64)
    np.savez_compressed(path(name), **d)


def load(name):
    z = np.load(path(name))
    out = []
    for i in range(int(z["n"])):
        f = dict(zip(SCALARS, [int(v) for v in z["s%d" % i]]))
        for k in KEYS:
            a = z["%s%d" % (k, i)]
            if k == "mvs":
                a = a.astype(np.int32)
            f[k] = a
        if "planesum%d" % i in z:
            f["planesum"] = int(z["planesum%d" % i][0])
        out.append(f)
    return out


def planesum(planes):
    return int(np.sum(planes.astype(np.uint64) * (np.arange(planes.size, dtype=np.uint64).reshape(planes.shape) % 251 + 1)))
