"""x265-amod_b200 -- B200-native lookahead for x265 (DJATOM/x265-aMod), Python binding.

The product is two native libraries built in-tree by __graft_entry__.build():
  lib/libx265cu.so   -- the CUDA engine behind the C ABI in include/x265cu.h
  lib/libx265la.so   -- the host Lookahead (host/lookahead.cpp) above that ABI, C surface la_capi.h
This module is a thin ctypes layer over libx265la.so for bench.py and the tests.  It fails
loudly if the libraries are missing or no CUDA device is usable -- there is no CPU fallback.
(The directory name carries a hyphen, so load it with importlib; see _pkg.py at the repo root.)
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# X265CU_LIBDIR: another build of the two libraries (kernel tuning variants, tools/build_variants.sh)
LIBDIR = os.environ.get("X265CU_LIBDIR") or os.path.join(HERE, "lib")

TYPE_AUTO, TYPE_IDR, TYPE_I, TYPE_P, TYPE_BREF, TYPE_B = 0, 1, 2, 3, 4, 5
TYPE_NAMES = {0: "AUTO", 1: "IDR", 2: "I", 3: "P", 4: "b", 5: "B"}


class Geometry(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("low_width", "low_height", "bw", "bh", "ncu", "stride", "plane_lines",
                                          "margin_x", "margin_y", "nb", "n_mv_stores", "n_cost_stores", "ncu_full")]


class LaParam(C.Structure):
    _fields_ = [("sourceWidth", C.c_int32), ("sourceHeight", C.c_int32), ("internalBitDepth", C.c_int32),
                ("maxCUSize", C.c_int32), ("fpsNum", C.c_int32), ("fpsDenom", C.c_int32),
                ("bframes", C.c_int32), ("lookaheadDepth", C.c_int32), ("bFrameAdaptive", C.c_int32),
                ("bBPyramid", C.c_int32), ("bFrameBias", C.c_int32), ("scenecutThreshold", C.c_int32),
                ("scenecutBias", C.c_double), ("keyframeMax", C.c_int32), ("keyframeMin", C.c_int32),
                ("bOpenGOP", C.c_int32), ("bIntraRefresh", C.c_int32), ("bEnableWeightedPred", C.c_int32),
                ("bEnableWeightedBiPred", C.c_int32), ("lookaheadSlices", C.c_int32), ("maxNumReferences", C.c_int32),
                ("aqMode", C.c_int32), ("aqStrength", C.c_double), ("cuTree", C.c_int32), ("qCompress", C.c_double),
                ("qgSize", C.c_int32), ("vbvBufferSize", C.c_int32), ("vbvMaxBitrate", C.c_int32),
                ("rateControlMode", C.c_int32), ("poolWorkers", C.c_int32), ("device", C.c_int32),
                ("extraSlots", C.c_int32), ("speculate", C.c_int32), ("pinHost", C.c_int32),
                ("asyncDepth", C.c_int32), ("pendingMax", C.c_int32), ("shardCount", C.c_int32), ("batchMin", C.c_int32), ("gopLookahead", C.c_int32), ("radl", C.c_int32),
                ("csvLogLevel", C.c_int32), ("numRowsPerSlice", C.c_int32), ("bEnableFades", C.c_int32),
                ("bEnableTemporalSubLayers", C.c_int32), ("bHistBasedSceneCut", C.c_int32),
                ("bEnableHME", C.c_int32), ("hmeSearchMethod", C.c_int32 * 2), ("hmeRange", C.c_int32 * 2)]


class FrameInfo(C.Structure):
    _fields_ = [("poc", C.c_int32), ("sliceType", C.c_int32), ("bScenecut", C.c_int32), ("bKeyframe", C.c_int32),
                ("bLastMiniGopBFrame", C.c_int32), ("leadingBframes", C.c_int32),
                ("pts", C.c_int64), ("reorderedPts", C.c_int64), ("satdCost", C.c_int64), ("handle", C.c_void_p),
                ("gopOffset", C.c_int32), ("gopId", C.c_int32), ("tempLayer", C.c_int32), ("gopIdWritten", C.c_int32)]


class FrameOut(C.Structure):
    _fields_ = [("intra_cost", C.c_void_p), ("intra_mode", C.c_void_p), ("qp_aq_offset", C.c_void_p),
                ("qp_cutree_offset", C.c_void_p), ("inv_qscale_factor", C.c_void_p), ("propagate_cost", C.c_void_p),
                ("planes", C.c_void_p), ("lowres_costs00", C.c_void_p), ("row_satds00", C.c_void_p)]


class Mirror(C.Structure):
    """x265la_mirror (host/la_capi.h): destinations of the asynchronous host mirror of a decided frame"""
    _fields_ = [("intraCost", C.c_void_p), ("qpAqOffset", C.c_void_p), ("qpCuTreeOffset", C.c_void_p),
                ("invQscaleFactor", C.c_void_p), ("planes", C.c_void_p), ("lowresMvs", (C.c_void_p * 18) * 2),
                ("d0", C.c_int32), ("d1", C.c_int32), ("lowresCosts", C.c_void_p), ("rowSatds", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("search_jobs", C.c_uint64), ("cost_jobs", C.c_uint64)]


# x265cu_exchange_fn (include/x265cu.h): user, bufs[nranks], bytes[nranks], nranks, cuda stream
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.c_int32, C.c_void_p)

K_NAMES = ["lowres", "aq", "intra", "search", "cost", "weight", "cutree"]

_libs = {}


def default_lib_path():
    return os.path.join(LIBDIR, "libx265la.so")


def load_lib(path=None):
    """Loads libx265la.so (and, through its DT_NEEDED/RPATH, libx265cu.so).  Raises if missing."""
    path = path or default_lib_path()
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError("native library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % path)
    lib = C.CDLL(path, mode=C.RTLD_LOCAL)
    lib.x265la_open.restype = C.c_void_p
    lib.x265la_open.argtypes = [C.POINTER(LaParam), C.c_char_p, C.c_int32]
    lib.x265la_close.argtypes = [C.c_void_p]
    lib.x265la_param_default.argtypes = [C.POINTER(LaParam)]
    lib.x265la_get_geometry.argtypes = [C.c_void_p, C.POINTER(Geometry)]
    lib.x265la_last_error.restype = C.c_char_p
    lib.x265la_last_error.argtypes = [C.c_void_p]
    lib.x265la_add_picture.restype = C.c_void_p
    lib.x265la_add_picture.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                       C.c_int64, C.c_int32, C.c_int32]
    lib.x265la_vbv_rows.argtypes = [C.c_void_p]
    lib.x265la_vbv_row_costs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 4
    lib.x265la_frame_mirror_async.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(Mirror), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)]
    lib.x265la_mirror_wait.argtypes = [C.c_void_p, C.c_int64]
    lib.x265la_pin.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.x265la_unpin.argtypes = [C.c_void_p, C.c_void_p]
    lib.x265la_estimated_picture_cost_dist.restype = C.c_int64
    lib.x265la_estimated_picture_cost_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.x265la_frame_planned.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    lib.x265la_flush.argtypes = [C.c_void_p]
    lib.x265la_get_decided.argtypes = [C.c_void_p, C.POINTER(FrameInfo)]
    lib.x265la_estimated_picture_cost.restype = C.c_int64
    lib.x265la_estimated_picture_cost.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.x265la_release.argtypes = [C.c_void_p, C.c_void_p]
    lib.x265la_frame_scalars.argtypes = [C.c_void_p] * 9
    lib.x265la_frame_mvs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.x265la_frame_hme_mvs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.x265la_frame_costs.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.x265la_frame_fetch.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(FrameOut)]
    lib.x265la_frame_weights.argtypes = [C.c_void_p] * 6
    lib.x265la_frame_fade.argtypes = [C.c_void_p] * 4
    lib.x265la_frame_hist.argtypes = [C.c_void_p] * 5
    lib.x265la_get_timers.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_int32]
    lib.x265la_engine.restype = C.c_void_p
    lib.x265la_engine.argtypes = [C.c_void_p]
    lib.x265la_shard_config.argtypes = [C.c_void_p, C.c_int32, C.c_int32, EXCHANGE_FN, C.c_void_p]
    _libs[path] = lib
    return lib


def load_engine(path=None):
    """ctypes handle on libx265cu.so itself (C ABI of include/x265cu.h)."""
    path = path or os.path.join(LIBDIR, "libx265cu.so")
    if not os.path.exists(path):
        raise RuntimeError("native library %s is missing (no CPU fallback)" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.x265cu_strerror.restype = C.c_char_p
    lib.x265cu_last_error.restype = C.c_char_p
    lib.x265cu_last_error.argtypes = [C.c_void_p]
    lib.x265cu_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    lib.x265cu_profile_enable.argtypes = [C.c_void_p, C.c_int32]
    lib.x265cu_profile_get.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int32]
    lib.x265cu_profile_get_busy.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.x265cu_sync.argtypes = [C.c_void_p]
    lib.x265cu_timer_start.argtypes = [C.c_void_p]
    lib.x265cu_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    lib.x265cu_pin_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    lib.x265cu_unpin_host.argtypes = [C.c_void_p, C.c_void_p]
    return lib


def make_param(width, height, depth=8, **kw):
    lib_defaults = LaParam()
    # defaults = x265 preset medium (source/common/param.cpp:164-349)
    d = dict(internalBitDepth=depth, maxCUSize=64, fpsNum=30, fpsDenom=1, bframes=4, lookaheadDepth=20,
             bFrameAdaptive=2, bBPyramid=1, bFrameBias=0, scenecutThreshold=40, scenecutBias=5.0,
             keyframeMax=250, keyframeMin=0, bOpenGOP=1, bIntraRefresh=0, bEnableWeightedPred=1,
             bEnableWeightedBiPred=0, lookaheadSlices=0, maxNumReferences=3, aqMode=2, aqStrength=1.0,
             cuTree=1, qCompress=0.6, qgSize=32, vbvBufferSize=0, vbvMaxBitrate=0, rateControlMode=2,
             poolWorkers=0, device=0, extraSlots=8, speculate=1, pinHost=0, asyncDepth=0, pendingMax=0, shardCount=0, batchMin=0, gopLookahead=0, radl=0, csvLogLevel=0, numRowsPerSlice=0, bEnableFades=0, bEnableTemporalSubLayers=0, bHistBasedSceneCut=0,
             bEnableHME=0, hmeSearchMethod=(1, 2), hmeRange=(16, 32))
    d.update(kw)
    lib_defaults.sourceWidth, lib_defaults.sourceHeight = width, height
    for k, v in d.items():
        if isinstance(v, (tuple, list)):
            v = (C.c_int32 * len(v))(*v)
        setattr(lib_defaults, k, v)
    return lib_defaults


class Lookahead:
    """Frames in (host numpy planes) -> decided frames out; mirrors the reference's Lookahead
    (addPicture / flush / getDecidedPicture / getEstimatedPictureCost)."""

    def __init__(self, width, height, depth=8, lib_path=None, **kw):
        self.lib = load_lib(lib_path)
        self.depth = depth
        self.param = make_param(width, height, depth, **kw)
        err = C.create_string_buffer(512)
        self.h = self.lib.x265la_open(C.byref(self.param), err, 512)
        if not self.h:
            raise RuntimeError("x265la_open failed: %s" % err.value.decode())
        self.geom = Geometry()
        self.lib.x265la_get_geometry(self.h, C.byref(self.geom))
        self._keep = {}      # handle -> planes kept alive until the frame is released
        self.dtype = np.uint8 if depth == 8 else np.uint16

    def add_picture_ptr(self, y_ptr, u_ptr, v_ptr, stride_y, stride_c, pts=0, slice_type=TYPE_AUTO, pass2_type=TYPE_AUTO):
        """addPicture with raw pointers (host, pinned host or device memory); the caller keeps them alive."""
        h = self.lib.x265la_add_picture(self.h, y_ptr, u_ptr, v_ptr, stride_y, stride_c, pts, pass2_type, slice_type)
        if not h:
            raise RuntimeError("addPicture failed: %s" % self.lib.x265la_last_error(self.h).decode())
        return h

    def add_picture(self, y, u, v, pts=0, slice_type=TYPE_AUTO, pass2_type=TYPE_AUTO):
        """slice_type: x265_picture::sliceType as an application forces it (-> Lowres::sliceTypeReq);
        pass2_type: the first-pass type of a 2-pass encode (the argument of the reference's Lookahead::addPicture)"""
        y = np.ascontiguousarray(y, self.dtype)
        u = None if u is None else np.ascontiguousarray(u, self.dtype)
        v = None if v is None else np.ascontiguousarray(v, self.dtype)
        h = self.lib.x265la_add_picture(self.h, y.ctypes.data, u.ctypes.data if u is not None else None,
                                        v.ctypes.data if v is not None else None, y.shape[1],
                                        u.shape[1] if u is not None else 0, pts, pass2_type, slice_type)
        if not h:
            raise RuntimeError("addPicture failed: %s" % self.lib.x265la_last_error(self.h).decode())
        self._keep[h] = (y, u, v)
        return h

    def flush(self):
        self.lib.x265la_flush(self.h)

    def get_decided(self):
        info = FrameInfo()
        r = self.lib.x265la_get_decided(self.h, C.byref(info))
        if r < 0:
            raise RuntimeError("getDecidedPicture failed: %s" % self.lib.x265la_last_error(self.h).decode())
        return info if r == 1 else None

    def estimated_picture_cost(self, frame, ref0=None, ref1=None):
        return self.lib.x265la_estimated_picture_cost(self.h, frame, ref0, ref1)

    def estimate_dict(self, frame, ref0=None, ref1=None, pir=(-1, -1), vbv=True):
        """getEstimatedPictureCost incl. its VBV half, in the layout of oracle/refbind.py's frame["est"]"""
        g = self.geom
        satd = self.estimated_picture_cost(frame, ref0, ref1)
        err = self.lib.x265la_last_error(self.h)
        if err:
            raise RuntimeError("getEstimatedPictureCost failed: %s" % err.decode())
        d = dict(satdCost=satd)
        if vbv:
            rows = self.lib.x265la_vbv_rows(self.h)
            a = np.zeros(rows, np.uint32); b = np.zeros(rows, np.uint32)
            rc = np.zeros(g.ncu, np.uint16); ic = np.zeros(g.ncu, np.int32)
            if self.lib.x265la_vbv_row_costs(self.h, frame, pir[0], pir[1], a.ctypes.data, b.ctypes.data, rc.ctypes.data, ic.ctypes.data) != 0:
                raise RuntimeError("x265la_vbv_row_costs failed: %s" % self.lib.x265la_last_error(self.h).decode())
            d.update(satdForVbv=a, intraSatdForVbv=b, lowresCostForRc=rc, intraCostForRc=ic)
        return d

    def release(self, handle):
        self._keep.pop(handle, None)
        self.lib.x265la_release(self.h, handle)

    def frame_dict(self, info, planes=False):
        """Published Lowres state of a decided frame in the same dict layout as oracle/refbind.py."""
        g, nb = self.geom, self.geom.nb
        ncu, bh, nfull = g.ncu, g.bh, g.ncu_full
        hnd = info.handle
        d = dict(poc=info.poc, sliceType=info.sliceType, bScenecut=info.bScenecut, bKeyframe=info.bKeyframe,
                 bLastMiniGopBFrame=info.bLastMiniGopBFrame, leadingBframes=info.leadingBframes,
                 bw=g.bw, bh=bh, nb=nb, stride=g.stride, planeLines=g.plane_lines)
        if self.param.bEnableTemporalSubLayers > 2:
            d.update(gopOffset=info.gopOffset, tempLayer=info.tempLayer, gopId=info.gopId if info.gopIdWritten else None,
                     gopIdTop=self.param.bEnableTemporalSubLayers - 3)
        costEst = np.zeros((nb, nb), np.int64); costEstAq = np.zeros((nb, nb), np.int64)
        intraMbs = np.zeros(nb, np.int32); valid = np.zeros((nb, nb), np.int32)
        ssd = np.zeros(3, np.uint64); sm = np.zeros(3, np.uint64); wd = np.zeros(nb, np.float64)
        self.lib.x265la_frame_scalars(self.h, hnd, costEst.ctypes.data, costEstAq.ctypes.data, intraMbs.ctypes.data,
                                      valid.ctypes.data, ssd.ctypes.data, sm.ctypes.data, wd.ctypes.data)
        d.update(costEst=costEst, costEstAq=costEstAq, intraMbs=intraMbs, rowSatdsValid=valid, wp_ssd=ssd, wp_sum=sm,
                 weightedCostDelta=wd)
        mvs = np.zeros((2, nb, ncu, 2), np.int32); mvc = np.zeros((2, nb, ncu), np.int32)
        searched = np.zeros((2, nb), bool)
        for l in range(2):
            for i in range(nb):
                searched[l, i] = bool(self.lib.x265la_frame_mvs(self.h, hnd, l, i, mvs[l, i].ctypes.data, mvc[l, i].ctypes.data))
        d.update(mvs=mvs, mvCosts=mvc, searched=searched)
        if self.param.bEnableHME and self.param.sourceHeight >= 540 and self.param.shardCount <= 1:
            # level-0 results of --hme behind every published search (Lowres::lowerResMvs / lowerResMvCosts)
            n4 = (((self.param.sourceWidth // 4) + 7) >> 3) * (((self.param.sourceHeight // 4) + 7) >> 3)
            lm = np.zeros((2, nb, n4, 2), np.int32); lmc = np.zeros((2, nb, n4), np.int32)
            for l in range(2):
                for i in range(nb):
                    if searched[l, i]:
                        self.lib.x265la_frame_hme_mvs(self.h, hnd, l, i, lm[l, i].ctypes.data, lmc[l, i].ctypes.data)
            d.update(lowerMvs=lm, lowerMvCosts=lmc)
        ws = np.zeros((4, nb), np.int32)
        self.lib.x265la_frame_weights(self.h, hnd, ws[0].ctypes.data, ws[1].ctypes.data, ws[2].ctypes.data, ws[3].ctypes.data)
        d.update(weightState=ws[0], weightParams=ws[1:].T.copy())
        fe = C.c_int32(0); fv = C.c_double(0)
        self.lib.x265la_frame_fade(self.h, hnd, C.byref(fe), C.byref(fv))
        d.update(bIsFadeEnd=fe.value, frameVariance=fv.value)
        hv = (C.c_int32 * 3)(); ha = (C.c_int32 * 3)(); hc = C.c_uint64(0)
        if self.lib.x265la_frame_hist(self.h, hnd, hv, ha, C.byref(hc)) == 0:
            d.update(histVar=list(hv), histAvg=list(ha), histCheck=int(hc.value))
        ps = np.zeros(251, np.int64); pt = np.zeros(251, np.int32); ib = C.c_int32(0)
        self.lib.x265la_frame_planned(self.h, hnd, ps.ctypes.data, pt.ctypes.data, 251, C.byref(ib))
        d.update(plannedSatd=ps, plannedType=pt, indB=ib.value)
        lc = np.zeros((nb, nb, ncu), np.uint16); rs = np.zeros((nb, nb, bh), np.int32)
        for i in range(nb):
            for j in range(nb):
                if valid[i, j]:
                    self.lib.x265la_frame_costs(self.h, hnd, i, j, lc[i, j].ctypes.data, rs[i, j].ctypes.data)
                else:
                    rs[i, j, 0] = -1
        d.update(lowresCosts=lc, rowSatds=rs)
        arr = dict(intraCost=np.zeros(ncu, np.int32), intraMode=np.zeros(ncu, np.uint8),
                   qpAqOffset=np.zeros(nfull, np.float64), qpCuTreeOffset=np.zeros(nfull, np.float64),
                   invQscaleFactor=np.zeros(nfull, np.int32), propagateCost=np.zeros(ncu, np.uint16))
        fo = FrameOut()
        fo.intra_cost = arr["intraCost"].ctypes.data; fo.intra_mode = arr["intraMode"].ctypes.data
        fo.qp_aq_offset = arr["qpAqOffset"].ctypes.data; fo.qp_cutree_offset = arr["qpCuTreeOffset"].ctypes.data
        fo.inv_qscale_factor = arr["invQscaleFactor"].ctypes.data; fo.propagate_cost = arr["propagateCost"].ctypes.data
        if planes:
            pl = np.zeros((4, g.plane_lines, g.stride), self.dtype)
            fo.planes = pl.ctypes.data
            arr["planes"] = pl
        if self.lib.x265la_frame_fetch(self.h, hnd, C.byref(fo)) != 0:
            raise RuntimeError("fetch failed: %s" % self.lib.x265la_last_error(self.h).decode())
        d.update(arr)
        return d

    def engine(self):
        return self.lib.x265la_engine(self.h)

    def shard(self, rank, nranks, exchange):
        """One stream over several ranks (open with shardCount=nranks): `exchange` is an EXCHANGE_FN, see shard.py."""
        self._exchange = exchange      # keep the ctypes thunk alive
        if self.lib.x265la_shard_config(self.h, rank, nranks, exchange, None) != 0:
            raise RuntimeError("x265la_shard_config failed: %s" % self.lib.x265la_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.x265la_close(self.h)
            self.h = None
            self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefTracker:
    """The nearest list-0 / list-1 reference of each decided frame, the way the encoder's DPB would hold them when
    Encoder::encode calls getEstimatedPictureCost (encoder.cpp:2367): frames arrive in coded order; references are the
    I / P / b-ref frames coded so far; an IDR empties the DPB first.  These are exactly the (p0, p1, b) estimates
    slicetypeDecide pre-computes for rate control (slicetype.cpp:2378-2427)."""

    def __init__(self):
        self.refs = []      # (poc, token) of reference frames coded so far

    def push(self, poc, slice_type, token):
        """returns (ref0_token, ref1_token) for this frame (None = no such reference), then files the frame"""
        if slice_type == TYPE_IDR:
            self.refs = []
        r0 = r1 = None
        if slice_type not in (TYPE_IDR, TYPE_I):
            lo = [r for r in self.refs if r[0] < poc]
            hi = [r for r in self.refs if r[0] > poc]
            if lo:
                r0 = max(lo)[1]
            if hi and slice_type in (TYPE_B, TYPE_BREF):
                r1 = min(hi)[1]
        if slice_type != TYPE_B:
            self.refs.append((poc, token))
            self.refs = self.refs[-4:]
        return r0, r1


def run_sequence(la, frames_iter, collect=True, planes=False, estimate_cost=False, slice_types=None, pass2_types=None,
                 pir=(-1, -1)):
    """Drive a Lookahead the way Encoder::encode does (one picture in, drain what is decided),
    then flush.  Returns the decided frames (dicts) in output order.  slice_types: {poc: forced type}
    (x265_picture::sliceType as an application may set it); pass2_types: {poc: first-pass type of a 2-pass encode}.
    estimate_cost: also call getEstimatedPictureCost (+ its VBV half) per decided frame like the encoder does, with the
    references RefTracker derives; results under the frame's "est" key."""
    out = []
    tracker = RefTracker()
    held = []       # handles still needed as references

    def drain():
        while True:
            info = la.get_decided()
            if info is None:
                break
            if collect:
                d = la.frame_dict(info, planes=planes)
            else:
                d = dict(poc=info.poc, sliceType=info.sliceType, bScenecut=info.bScenecut, bKeyframe=info.bKeyframe)
            if estimate_cost:
                r0, r1 = tracker.push(info.poc, info.sliceType, info.handle)
                d["est"] = la.estimate_dict(info.handle, r0, r1, pir=pir)
                if collect:
                    rs = np.zeros(la.geom.bh, np.int32)
                    la.lib.x265la_frame_costs(la.h, info.handle, (info.poc - la_poc(r0)) if r0 else 0,
                                              (la_poc(r1) - info.poc) if r1 else 0, None, rs.ctypes.data)
                    d["est"]["rowSatds"] = rs
                held.append((info.poc, info.handle))
                live = set(t for _, t in tracker.refs)
                for ph in list(held):
                    if ph[1] not in live and ph[1] != info.handle:
                        held.remove(ph)
                        la.release(ph[1])
                if info.sliceType == TYPE_B:
                    held.remove((info.poc, info.handle))
                    la.release(info.handle)
            else:
                la.release(info.handle)
            out.append(d)

    pocs = {}

    def la_poc(handle):
        return pocs[handle]

    _push = tracker.push

    def push(poc, t, token):
        pocs[token] = poc
        return _push(poc, t, token)
    tracker.push = push

    for i, (y, u, v) in enumerate(frames_iter):
        la.add_picture(y, u, v, pts=i, slice_type=(slice_types or {}).get(i, TYPE_AUTO),
                       pass2_type=(pass2_types or {}).get(i, TYPE_AUTO))
        drain()
    la.flush()
    drain()
    for _, hnd in held:
        la.release(hnd)
    return out
