"""ctypes binding to oracle/_ref/libx265ref{8,10}.so (the UNMODIFIED reference lookahead built
by oracle/Makefile.ref plus oracle/ref_harness.cpp).  TEST INFRASTRUCTURE ONLY: imported by
tests/, tools that regenerate tests/golden/, __graft_entry__.smoke() and bench.py's
reference / cpu_baseline legs -- never by the product path."""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class RefLaConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("fpsNum", C.c_int32), ("fpsDenom", C.c_int32),
                ("bframes", C.c_int32), ("lookaheadDepth", C.c_int32), ("bFrameAdaptive", C.c_int32),
                ("bBPyramid", C.c_int32), ("scenecutThreshold", C.c_int32), ("keyframeMax", C.c_int32),
                ("keyframeMin", C.c_int32), ("bOpenGOP", C.c_int32), ("aqMode", C.c_int32),
                ("aqStrength", C.c_double), ("cuTree", C.c_int32), ("qCompress", C.c_double),
                ("weightp", C.c_int32), ("weightb", C.c_int32), ("poolThreads", C.c_int32),
                ("lookaheadSlices", C.c_int32), ("qgSize", C.c_int32), ("bFrameBias", C.c_int32),
                ("scenecutBias", C.c_double), ("vbvBufferSize", C.c_int32), ("vbvMaxBitrate", C.c_int32),
                ("bitrate", C.c_int32), ("dumpPlanes", C.c_int32), ("bIntraRefresh", C.c_int32),
                ("gopLookahead", C.c_int32), ("radl", C.c_int32), ("keepFrames", C.c_int32),
                ("fades", C.c_int32), ("temporalLayers", C.c_int32), ("histScenecut", C.c_int32), ("csp400", C.c_int32),
                ("hme", C.c_int32), ("hmeSearch0", C.c_int32), ("hmeSearch1", C.c_int32), ("hmeRange0", C.c_int32),
                ("hmeRange1", C.c_int32)]


class RefLaFrame(C.Structure):
    _fields_ = [("poc", C.c_int32), ("sliceType", C.c_int32), ("bScenecut", C.c_int32), ("bKeyframe", C.c_int32),
                ("bLastMiniGopBFrame", C.c_int32), ("leadingBframes", C.c_int32),
                ("bw", C.c_int32), ("bh", C.c_int32), ("nb", C.c_int32), ("stride", C.c_int32),
                ("planeLines", C.c_int32), ("satdCost", C.c_int64),
                ("costEst", C.c_void_p), ("costEstAq", C.c_void_p), ("intraMbs", C.c_void_p),
                ("rowSatds", C.c_void_p), ("lowresCosts", C.c_void_p), ("mvs", C.c_void_p),
                ("mvCosts", C.c_void_p), ("intraCost", C.c_void_p), ("intraMode", C.c_void_p),
                ("qpAqOffset", C.c_void_p), ("qpCuTreeOffset", C.c_void_p), ("invQscaleFactor", C.c_void_p),
                ("propagateCost", C.c_void_p), ("wp_ssd", C.c_uint64 * 3), ("wp_sum", C.c_uint64 * 3),
                ("weightedCostDelta", C.c_void_p), ("planes", C.c_void_p),
                ("ncuFull", C.c_int32), ("indB", C.c_int32), ("plannedSatd", C.c_void_p), ("plannedType", C.c_void_p),
                ("estimated", C.c_int32), ("vbvRows", C.c_int32), ("estSatdCost", C.c_int64),
                ("satdForVbv", C.c_void_p), ("intraSatdForVbv", C.c_void_p), ("lowresCostForRc", C.c_void_p),
                ("intraCostForRc", C.c_void_p), ("estRowSatds", C.c_void_p),
                ("bIsFadeEnd", C.c_int32), ("pad0", C.c_int32), ("frameVariance", C.c_double),
                ("histVar", C.c_int32 * 3), ("histAvg", C.c_int32 * 3), ("histCheck", C.c_uint64),
                ("bw4", C.c_int32), ("bh4", C.c_int32), ("lowerMvs", C.c_void_p), ("lowerMvCosts", C.c_void_p),
                ("gopOffset", C.c_int32), ("gopId", C.c_int32), ("tempLayer", C.c_int32), ("pad1", C.c_int32)]


DEFAULTS = dict(fpsNum=30, fpsDenom=1, bframes=4, lookaheadDepth=20, bFrameAdaptive=2, bBPyramid=1,
                scenecutThreshold=40, keyframeMax=250, keyframeMin=0, bOpenGOP=1, aqMode=2, aqStrength=1.0,
                cuTree=1, qCompress=0.6, weightp=1, weightb=0, poolThreads=0, lookaheadSlices=0, qgSize=32,
                bFrameBias=0, scenecutBias=5.0, vbvBufferSize=0, vbvMaxBitrate=0, bitrate=0, dumpPlanes=0,
                bIntraRefresh=0, gopLookahead=0, radl=0, keepFrames=0, fades=0, temporalLayers=0, histScenecut=0, csp400=0,
                hme=0, hmeSearch0=1, hmeSearch1=2, hmeRange0=16, hmeRange1=32)


def lib_path(depth):
    return os.path.join(HERE, "_ref", "libx265ref%d.so" % (depth if depth in (8, 10, 12) else 10))


def available(depth=8):
    return os.path.exists(lib_path(depth))


_libs = {}


def load(depth):
    key = depth if depth in (8, 10, 12) else 10
    if key in _libs:
        return _libs[key]
    lib = C.CDLL(lib_path(depth), mode=C.RTLD_LOCAL)
    lib.ref_la_open.restype = C.c_void_p
    lib.ref_la_open.argtypes = [C.POINTER(RefLaConfig)]
    lib.ref_la_put.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.ref_la_put_typed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_la_snapshot_seconds.argtypes = [C.c_void_p]
    lib.ref_la_snapshot_seconds.restype = C.c_double
    lib.ref_la_drop.argtypes = [C.c_void_p, C.c_int]
    lib.ref_la_estimate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.ref_la_flush.argtypes = [C.c_void_p, C.c_int]
    lib.ref_la_num_out.argtypes = [C.c_void_p]
    lib.ref_la_seconds.argtypes = [C.c_void_p]
    lib.ref_la_seconds.restype = C.c_double
    lib.ref_la_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(RefLaFrame)]
    lib.ref_la_close.argtypes = [C.c_void_p]
    lib.ref_la_effective.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    lib.ref_mvcost_table.argtypes = [C.c_void_p, C.c_int]
    lib.ref_satd8x8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.ref_sad8x8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.ref_exp2fix8.argtypes = [C.c_double]
    lib.ref_simd_install.argtypes = [C.c_int]
    lib.ref_simd_selftest.argtypes = [C.c_int]
    lib.ref_setup_primitives()
    _libs[key] = lib
    return lib


def _arr(ptr, dtype, n):
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def make_config(width, height, **kw):
    cfg = RefLaConfig()
    d = dict(DEFAULTS)
    d.update(kw)
    cfg.width, cfg.height = width, height
    for k, v in d.items():
        setattr(cfg, k, v)
    return cfg


class RefLookahead:
    """The reference's Lookahead, frames in -> decided frames (full Lowres snapshots) out."""

    def __init__(self, width, height, depth=8, **kw):
        self.depth = depth
        self.lib = load(depth)
        self.cfg = make_config(width, height, **kw)
        self.h = self.lib.ref_la_open(C.byref(self.cfg))
        if not self.h:
            raise RuntimeError("reference encoder_open failed")
        eff = (C.c_int32 * 16)()
        self.lib.ref_la_effective(self.h, eff)
        names = ["width", "height", "keyframeMin", "keyframeMax", "bframes", "lookaheadDepth", "bFrameAdaptive",
                 "bBPyramid", "aqMode", "cuTree", "qgSize", "lookaheadSlices", "poolWorkers", "weightp", "weightb",
                 "scenecutThreshold"]
        self.effective = dict(zip(names, list(eff)))
        self._fetched = 0

    def put(self, y, u, v, snap=True, slice_type=0, pass2_type=0):
        """slice_type: x265_picture::sliceType forced by the application (-> Lowres::sliceTypeReq);
        pass2_type: the first-pass type a 2-pass encode hands to Lookahead::addPicture"""
        dt = np.uint8 if self.depth == 8 else np.uint16
        y = np.ascontiguousarray(y, dt); u = np.ascontiguousarray(u, dt); v = np.ascontiguousarray(v, dt)
        return self.lib.ref_la_put_typed(self.h, y.ctypes.data, u.ctypes.data, v.ctypes.data,
                                         y.shape[1], u.shape[1], 1 if snap else 0, int(pass2_type), int(slice_type))

    def num_out(self):
        return self.lib.ref_la_num_out(self.h)

    def out_info(self, idx):
        """(poc, sliceType) of decided frame idx (output order)"""
        f = RefLaFrame()
        self.lib.ref_la_get(self.h, idx, C.byref(f))
        return f.poc, f.sliceType

    def estimate(self, idx, ref0=-1, ref1=-1, pir=(0, 0)):
        """Lookahead::getEstimatedPictureCost on decided frame idx with references ref0 / ref1 (output indices)"""
        if self.lib.ref_la_estimate(self.h, idx, ref0, ref1, pir[0], pir[1]) != 0:
            raise RuntimeError("ref_la_estimate: frame %d or its references are no longer alive" % idx)

    def frame(self, idx, drop=False):
        """decided frame idx as a dict; drop=True frees the harness' copy afterwards (full-size runs)"""
        f = RefLaFrame()
        self.lib.ref_la_get(self.h, idx, C.byref(f))
        d = self._to_dict(f)
        if drop:
            self.lib.ref_la_drop(self.h, idx)
        return d

    def snapshot_seconds(self):
        return self.lib.ref_la_snapshot_seconds(self.h)

    def flush(self, snap=True):
        return self.lib.ref_la_flush(self.h, 1 if snap else 0)

    def seconds(self):
        return self.lib.ref_la_seconds(self.h)

    def frames(self):
        """All decided frames so far, in output (coded) order, as dicts of numpy arrays."""
        out = []
        n = self.lib.ref_la_num_out(self.h)
        for i in range(n):
            f = RefLaFrame()
            self.lib.ref_la_get(self.h, i, C.byref(f))
            out.append(self._to_dict(f))
        return out

    def _to_dict(self, f):
        nb, bw, bh = f.nb, f.bw, f.bh
        ncu = bw * bh
        d = dict(poc=f.poc, sliceType=f.sliceType, bScenecut=f.bScenecut, bKeyframe=f.bKeyframe,
                 bLastMiniGopBFrame=f.bLastMiniGopBFrame, leadingBframes=f.leadingBframes,
                 bw=bw, bh=bh, nb=nb, stride=f.stride, planeLines=f.planeLines, satdCost=f.satdCost,
                 wp_ssd=np.array(list(f.wp_ssd), np.uint64), wp_sum=np.array(list(f.wp_sum), np.uint64),
                 bIsFadeEnd=f.bIsFadeEnd, frameVariance=f.frameVariance,
                 histVar=list(f.histVar), histAvg=list(f.histAvg), histCheck=int(f.histCheck))
        if self.cfg.temporalLayers > 2:     # otherwise nothing ever writes them (Frame's constructor leaves m_gopOffset / m_gopId alone)
            d.update(gopOffset=f.gopOffset, gopId=f.gopId, tempLayer=f.tempLayer)
        if not nb:
            return d
        d["costEst"] = _arr(f.costEst, np.int64, nb * nb).reshape(nb, nb)
        d["costEstAq"] = _arr(f.costEstAq, np.int64, nb * nb).reshape(nb, nb)
        d["intraMbs"] = _arr(f.intraMbs, np.int32, nb)
        d["rowSatds"] = _arr(f.rowSatds, np.int32, nb * nb * bh).reshape(nb, nb, bh)
        d["lowresCosts"] = _arr(f.lowresCosts, np.uint16, nb * nb * ncu).reshape(nb, nb, ncu)
        d["mvs"] = _arr(f.mvs, np.int32, 2 * nb * ncu * 2).reshape(2, nb, ncu, 2)
        d["mvCosts"] = _arr(f.mvCosts, np.int32, 2 * nb * ncu).reshape(2, nb, ncu)
        d["intraCost"] = _arr(f.intraCost, np.int32, ncu)
        d["intraMode"] = _arr(f.intraMode, np.uint8, ncu)
        nfull = f.ncuFull or ncu
        d["qpAqOffset"] = _arr(f.qpAqOffset, np.float64, nfull)
        d["qpCuTreeOffset"] = _arr(f.qpCuTreeOffset, np.float64, nfull)
        d["invQscaleFactor"] = _arr(f.invQscaleFactor, np.int32, nfull)
        d["indB"] = f.indB
        d["plannedSatd"] = _arr(f.plannedSatd, np.int64, 251)
        d["plannedType"] = _arr(f.plannedType, np.int32, 251)
        if f.estimated:
            d["est"] = dict(satdCost=f.estSatdCost, satdForVbv=_arr(f.satdForVbv, np.uint32, f.vbvRows),
                            intraSatdForVbv=_arr(f.intraSatdForVbv, np.uint32, f.vbvRows),
                            lowresCostForRc=_arr(f.lowresCostForRc, np.uint16, ncu),
                            intraCostForRc=_arr(f.intraCostForRc, np.int32, ncu),
                            rowSatds=_arr(f.estRowSatds, np.int32, bh))
        d["propagateCost"] = _arr(f.propagateCost, np.uint16, ncu)
        d["weightedCostDelta"] = _arr(f.weightedCostDelta, np.float64, nb)
        if f.lowerMvs:
            n4 = f.bw4 * f.bh4
            d["bw4"], d["bh4"] = f.bw4, f.bh4
            d["lowerMvs"] = _arr(f.lowerMvs, np.int32, 2 * nb * n4 * 2).reshape(2, nb, n4, 2)
            d["lowerMvCosts"] = _arr(f.lowerMvCosts, np.int32, 2 * nb * n4).reshape(2, nb, n4)
        if f.planes:
            dt = np.uint8 if self.depth == 8 else np.uint16
            d["planes"] = _arr(f.planes, dt, 4 * f.stride * f.planeLines).reshape(4, f.planeLines, f.stride)
        return d

    def close(self):
        if self.h:
            self.lib.ref_la_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def simd_install(depth, on):
    """install (on=True) / remove the SSE4.1 intrinsics shims (oracle/ref_simd.cpp) in the reference's primitives table;
    returns True when they are active"""
    return bool(load(depth).ref_simd_install(1 if on else 0))


def simd_selftest(depth, iterations=3000):
    """mismatches of the shims against the C primitives on the pixelharness buffer recipe (0 = bit-exact, -1 = no SSE4.1)"""
    return load(depth).ref_simd_selftest(iterations)


def mvcost_table(depth, n=4096):
    lib = load(depth)
    out = np.zeros(2 * n + 1, np.uint16)
    qp = lib.ref_mvcost_table(out.ctypes.data, n)
    return qp, out
