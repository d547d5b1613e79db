/* slicetype_cuda.cpp -- the ENABLE_CUDA body of the reference's `class Lookahead`.
 *
 * Compiled INTO the reference encoder (DJATOM/x265-aMod 3.6+1) when it is built with -DENABLE_CUDA=1, next to the patched
 * source/encoder/slicetype.cpp (integration/x265_enable_cuda.patch): every public entry point of `Lookahead`
 * (encoder/slicetype.h:214-227, unchanged) forwards to one of the functions below, which drive the B200 lookahead through
 * its C surface (x265-amod_b200/host/la_capi.h -> include/x265cu.h).  The class declaration, its queues
 * (m_inputQueue / m_outputQueue / m_filled, which Encoder reads directly, encoder.cpp:1505,1661,1867,2541-2549) and every
 * caller in the encoder stay as they are: source/encoder/slicetype.h, encoder.cpp, ratecontrol.cpp, frameencoder.cpp,
 * search.cpp, weightPrediction.cpp are NOT touched.
 *
 * What happens where:
 *   addPicture             the picture's PicYuv planes go to the GPU (async H2D from the page-locked PicYuv buffer), the
 *                          non-pixel half of Lowres::init (common/lowres.cpp:337-365) runs here; nothing is computed on the CPU.
 *   getDecidedPicture      decisions come from the host library; everything the main encoder reads of the frame's Lowres is
 *                          mirrored into the host arrays Lowres::create allocated (SURVEY 8b "output contract"):
 *                          slice type and flags, costEst / costEstAq / intraMbs, qpAqOffset / qpCuTreeOffset /
 *                          invQscaleFactor, intraCost, wp_ssd / wp_sum, plannedSatd / plannedType / indB, every published
 *                          lowresMvs list (sentinel 0x7FFF for the others, search.cpp:1975-1977), and -- with weightp /
 *                          weightb -- the four lowres planes weightPrediction.cpp:60-88,354-365 reads.
 *   getEstimatedPictureCost  satdCost, and with VBV the row sums / rescaled lowresCostForRc / intraCost of slicetype.cpp:
 *                          1387-1436 computed on the GPU and written where the reference writes them.
 * There is no CPU fallback: if the GPU library cannot be opened Lookahead::create() fails and the encoder aborts.
 *
 * This file contains no reference source text; it only uses the reference's headers. */
#if ENABLE_CUDA

#include "common.h"
#include "frame.h"
#include "framedata.h"
#include "picyuv.h"
#include "lowres.h"
#include "slice.h"
#include "threadpool.h"
#include "slicetype.h"
#include "slicetype_cuda.h"

#include "la_capi.h"

#include <map>
#include <vector>
#include <stdlib.h>
#include <string.h>

namespace X265_NS {

namespace {

struct CudaState
{
    void* la;
    x265cu_geometry geom;
    std::map<void*, Frame*> frameOf;        /* library frame handle -> encoder Frame */
    std::map<Frame*, void*> handleOf;
    std::vector<void*> unreleased;          /* decided frames whose slot the library still holds for us */
    std::map<Frame*, int64_t> mirrorTicket; /* asynchronous mirror of the frame's Lowres still in flight */
    std::vector<void*> pinned;              /* page-locked Lowres arrays (once per Frame; frames are recycled by the DPB) */
    std::map<Frame*, bool> pinnedFrame;
    bool weightPlanesP, weightPlanesB;
};

Lock g_stateLock;
std::map<const Lookahead*, CudaState*> g_state;

CudaState* stateOf(const Lookahead& self)
{
    ScopedLock lock(g_stateLock);
    std::map<const Lookahead*, CudaState*>::iterator it = g_state.find(&self);
    return it == g_state.end() ? NULL : it->second;
}

int envInt(const char* name, int dflt)
{
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

void pin(CudaState* st, void* p, size_t bytes)
{
    if (p && x265la_pin(st->la, p, bytes) == 0) st->pinned.push_back(p);
}

/* page-lock the arrays of f's Lowres the mirror writes, so that the copies are truly asynchronous */
void pinLowres(CudaState* st, Frame* f, int nb)
{
    if (st->pinnedFrame.count(f)) return;
    st->pinnedFrame[f] = true;
    Lowres& l = f->m_lowres;
    const size_t ncu = st->geom.ncu, nfull = st->geom.ncu_full;
    pin(st, l.buffer[0], (size_t)4 * st->geom.stride * st->geom.plane_lines * sizeof(pixel));
    pin(st, l.intraCost, ncu * sizeof(int32_t));
    if (l.qpAqOffset && l.invQscaleFactor)
    {
        pin(st, l.qpAqOffset, nfull * sizeof(double)); pin(st, l.qpCuTreeOffset, nfull * sizeof(double));
        pin(st, l.invQscaleFactor, nfull * sizeof(int));
    }
    for (int list = 0; list < 2; list++)
        for (int d = 1; d < nb; d++)
            pin(st, l.lowresMvs[list][d], ncu * sizeof(MV));
}

void waitMirror(CudaState* st, Frame* f)
{
    std::map<Frame*, int64_t>::iterator it = st->mirrorTicket.find(f);
    if (it == st->mirrorTicket.end()) return;
    x265la_mirror_wait(st->la, it->second);
    st->mirrorTicket.erase(it);
}

void releaseHandle(CudaState* st, void* h)
{
    for (size_t i = 0; i < st->unreleased.size(); i++)
        if (st->unreleased[i] == h)
        {
            st->unreleased.erase(st->unreleased.begin() + i);
            x265la_release(st->la, h);
            return;
        }
}

} // namespace

bool cudaLookaheadCreate(Lookahead& self)
{
    x265_param* p = self.m_param;
    const char* why = NULL;
    if (p->bEnableHME && (p->hmeSearchMethod[0] == X265_SEA || p->hmeSearchMethod[1] == X265_SEA)) why = "--hme-search sea at levels 0 and 1";
    else if (p->bHistBasedSceneCut && X265_DEPTH != 8) why = "--hist-scenecut at high bit depth";
    else if (p->rc.aqMode > X265_AQ_AUTO_VARIANCE_BIASED && p->recursionSkipMode == EDGE_BASED_RSKIP) why = "--aq-mode 4/5 with --rskip 2 (the encoder reads the lookahead's full-resolution edge picture)";
    else if (p->rc.aqMode > X265_AQ_AUTO_VARIANCE_BIASED && p->bEnableFades) why = "--aq-mode 4/5 with --fades";
    else if (p->rc.hevcAq) why = "--hevc-aq";
    else if (p->bAQMotion) why = "--aq-motion";
    else if (p->analysisLoad || p->bAnalysisType == AVC_INFO) why = "--analysis-load";
    else if (p->bEnableFades && p->rc.qgSize == 8) why = "--fades with --qg-size 8";
    else if (p->bDynamicRefine) why = "--dynamic-refine";
    else if (p->rc.bStatRead && p->rc.cuTree) why = "2-pass cutree";
    else if (p->internalCsp != X265_CSP_I420 && p->internalCsp != X265_CSP_I400) why = "chroma formats other than 4:2:0 / 4:0:0";
    if (why)
    {
        x265_log(p, X265_LOG_ERROR, "ENABLE_CUDA lookahead: %s is not supported by the GPU path (no CPU fallback)\n", why);
        return false;
    }
    x265la_param q;
    x265la_param_default(&q);
    q.sourceWidth = p->sourceWidth; q.sourceHeight = p->sourceHeight; q.internalBitDepth = X265_DEPTH; q.maxCUSize = p->maxCUSize;
    q.fpsNum = p->fpsNum; q.fpsDenom = p->fpsDenom;
    q.bframes = p->bframes; q.lookaheadDepth = p->lookaheadDepth; q.bFrameAdaptive = p->bFrameAdaptive;
    q.bBPyramid = p->bBPyramid; q.bFrameBias = p->bFrameBias;
    q.scenecutThreshold = p->scenecutThreshold; q.scenecutBias = p->scenecutBias * 100.0;  /* configure() already divided by 100 */
    q.keyframeMax = p->keyframeMax; q.keyframeMin = p->keyframeMin; q.bOpenGOP = p->bOpenGOP; q.bIntraRefresh = p->bIntraRefresh;
    q.bEnableWeightedPred = p->bEnableWeightedPred; q.bEnableWeightedBiPred = p->bEnableWeightedBiPred;
    q.lookaheadSlices = p->lookaheadSlices;
    q.numRowsPerSlice = self.m_numCoopSlices > 1 ? self.m_numRowsPerSlice : 0;
    q.maxNumReferences = p->maxNumReferences;
    q.aqMode = p->rc.aqMode; q.aqStrength = p->rc.aqStrength; q.cuTree = p->rc.cuTree; q.qCompress = p->rc.qCompress;
    q.qgSize = p->rc.qgSize; q.vbvBufferSize = p->rc.vbvBufferSize; q.vbvMaxBitrate = p->rc.vbvMaxBitrate;
    q.rateControlMode = p->rc.rateControlMode;
    q.poolWorkers = self.m_pool ? self.m_pool->m_numWorkers : 0;
    q.gopLookahead = p->gopLookahead; q.radl = p->radl; q.csvLogLevel = p->csvLogLevel;
    q.bEnableFades = p->bEnableFades; q.bEnableTemporalSubLayers = p->bEnableTemporalSubLayers;
    q.bHistBasedSceneCut = p->bHistBasedSceneCut;
    q.bEnableHME = p->bEnableHME;
    for (int i = 0; i < 2; i++) { q.hmeSearchMethod[i] = p->hmeSearchMethod[i]; q.hmeRange[i] = p->hmeRange[i]; }
    q.device = envInt("X265_CUDA_DEVICE", 0);
    /* extra frames of input delay that keep the GPU busy while the host decides (same decisions, LookaheadParam::asyncDepth) */
    q.asyncDepth = envInt("X265_CUDA_ASYNC_DEPTH", 16);
    q.speculate = 1;
    q.pinHost = 1;                  /* PicYuv buffers are recycled through the DPB free list: page-lock each once */
    q.extraSlots = 8 + 2 * p->frameNumThreads;
    char err[512];
    err[0] = 0;
    void* la = x265la_open(&q, err, sizeof(err));
    if (!la)
    {
        x265_log(p, X265_LOG_ERROR, "ENABLE_CUDA lookahead: %s\n", err);
        return false;
    }
    CudaState* st = new CudaState;
    st->la = la;
    x265la_get_geometry(la, &st->geom);
    st->weightPlanesP = !!p->bEnableWeightedPred;
    st->weightPlanesB = !!p->bEnableWeightedBiPred;
    {
        ScopedLock lock(g_stateLock);
        g_state[&self] = st;
    }
    x265_log(p, X265_LOG_INFO, "lookahead on CUDA device %d (B200 engine, async depth %d, pool emulation %d workers)\n",
             q.device, q.asyncDepth, q.poolWorkers);
    return true;
}

void cudaLookaheadDestroy(Lookahead& self)
{
    CudaState* st = NULL;
    {
        ScopedLock lock(g_stateLock);
        std::map<const Lookahead*, CudaState*>::iterator it = g_state.find(&self);
        if (it != g_state.end()) { st = it->second; g_state.erase(it); }
    }
    if (!st) return;
    for (std::map<Frame*, int64_t>::iterator it = st->mirrorTicket.begin(); it != st->mirrorTicket.end(); ++it)
        x265la_mirror_wait(st->la, it->second);
    for (size_t i = 0; i < st->pinned.size(); i++)
        x265la_unpin(st->la, st->pinned[i]);
    x265la_close(st->la);
    delete st;
}

/* Lookahead::addPicture (slicetype.cpp:1200-1229) + the non-pixel half of Lowres::init (lowres.cpp:337-365) */
void cudaLookaheadAddPicture(Lookahead& self, Frame& f, int sliceType)
{
    CudaState* st = stateOf(self);
    x265_param* p = self.m_param;
    /* checkLookaheadQueue's fill rule; the library applies the same one (plus its async depth) to decide when frames flow */
    Lowres& l = f.m_lowres;
    l.sliceType = sliceType;
    l.bLastMiniGopBFrame = false; l.bKeyframe = false; l.bIsFadeEnd = false;
    l.frameNum = f.m_poc; l.leadingBframes = 0; l.indB = 0;
    memset(l.costEst, -1, sizeof(l.costEst));
    memset(l.weightedCostDelta, 0, sizeof(l.weightedCostDelta));
    if (l.qpAqOffset && l.invQscaleFactor)
        memset(l.costEstAq, -1, sizeof(l.costEstAq));
    for (int y = 0; y < p->bframes + 2; y++)
        for (int x = 0; x < p->bframes + 2; x++)
            l.rowSatds[y][x][0] = -1;
    for (int i = 0; i < p->bframes + 2; i++)
    {
        l.lowresMvs[0][i][0].x = 0x7FFF;
        l.lowresMvs[1][i][0].x = 0x7FFF;
        l.intraMbs[i] = 0;
    }
    if (p->rc.vbvBufferSize)
        for (int i = 0; i < X265_LOOKAHEAD_MAX + 1; i++)
            l.plannedType[i] = X265_TYPE_AUTO;
    l.fpelPlane[0] = l.lowresPlane[0];
    f.m_lowresInit = true;

    PicYuv* pic = f.m_fencPic;
    const bool chroma = pic->m_picCsp != X265_CSP_I400;
    void* h = x265la_add_picture(st->la, pic->m_picOrg[0], chroma ? pic->m_picOrg[1] : NULL, chroma ? pic->m_picOrg[2] : NULL,
                                 (int32_t)pic->m_stride, (int32_t)pic->m_strideC, f.m_pts, sliceType, l.sliceTypeReq);
    if (!h)
    {
        x265_log(p, X265_LOG_ERROR, "ENABLE_CUDA lookahead: addPicture failed: %s\n", x265la_last_error(st->la));
        return;
    }
    st->frameOf[h] = &f;
    st->handleOf[&f] = h;
    self.m_inputLock.acquire();
    self.m_inputQueue.pushBack(f);
    self.m_inputLock.release();
    self.m_inputCount++;
    /* Encoder::encode gates on m_filled only through getDecidedPicture; keep the reference's flag meaningful */
    if (!self.m_filled)
    {
        if (!p->bframes & !p->lookaheadDepth)
            self.m_filled = true;
        else if (self.m_inputCount >= p->lookaheadDepth + 2 + p->bframes)
            self.m_filled = true;
    }
}

void cudaLookaheadFlush(Lookahead& self)
{
    CudaState* st = stateOf(self);
    if (st) x265la_flush(st->la);
}

/* Lookahead::getDecidedPicture (slicetype.cpp:1289-1322) + the host mirror of the decided frame's Lowres */
Frame* cudaLookaheadGetDecided(Lookahead& self)
{
    CudaState* st = stateOf(self);
    if (!st || !self.m_filled) return NULL;
    x265_param* p = self.m_param;
    x265la_frame_info info;
    const int r = x265la_get_decided(st->la, &info);
    if (r < 0)
    {
        x265_log(p, X265_LOG_ERROR, "ENABLE_CUDA lookahead: %s\n", x265la_last_error(st->la));
        return NULL;
    }
    if (r == 0) return NULL;
    std::map<void*, Frame*>::iterator it = st->frameOf.find(info.handle);
    if (it == st->frameOf.end()) return NULL;
    Frame* f = it->second;
    st->frameOf.erase(it);
    Lowres& l = f->m_lowres;
    const int nb = st->geom.nb, ncu = st->geom.ncu;

    l.sliceType = info.sliceType; l.bScenecut = !!info.bScenecut; l.bKeyframe = !!info.bKeyframe;
    l.bLastMiniGopBFrame = !!info.bLastMiniGopBFrame; l.leadingBframes = info.leadingBframes;
    l.ipCostRatio = x265la_frame_ip_cost_ratio(st->la, info.handle);
    if (p->bEnableFades)
    {
        int32_t fadeEnd = 0; double variance = 0;
        x265la_frame_fade(st->la, info.handle, &fadeEnd, &variance);
        l.bIsFadeEnd = !!fadeEnd; l.frameVariance = variance;     /* ratecontrol.cpp:1416 reads bIsFadeEnd */
    }
    f->m_reorderedPts = info.reorderedPts;
    if (p->bEnableTemporalSubLayers > 2)
    {
        /* where in which random-access structure the frame sits, and its temporal layer (slicetype.cpp:2133-2320; the DPB
         * derives the reference picture sets from them) */
        f->m_gopOffset = info.gopOffset; f->m_tempLayer = (int8_t)info.tempLayer;
        if (info.gopIdWritten) f->m_gopId = (int8_t)info.gopId;
    }

    std::vector<int64_t> ce(nb * nb), cea(nb * nb);
    std::vector<int32_t> mbs(nb), valid(nb * nb);
    std::vector<double> wd(nb);
    x265la_frame_scalars(st->la, info.handle, &ce[0], &cea[0], &mbs[0], &valid[0], l.wp_ssd, l.wp_sum, &wd[0]);
    for (int i = 0; i < nb; i++)
    {
        for (int j = 0; j < nb; j++)
        {
            l.costEst[i][j] = ce[i * nb + j];
            if (l.qpAqOffset && l.invQscaleFactor) l.costEstAq[i][j] = cea[i * nb + j];
        }
        l.intraMbs[i] = mbs[i];
        l.weightedCostDelta[i] = wd[i];
    }
    if (p->rc.vbvBufferSize)
    {
        int32_t indB = 0;
        std::vector<int32_t> pt(X265_LOOKAHEAD_MAX + 1);
        x265la_frame_planned(st->la, info.handle, l.plannedSatd, &pt[0], X265_LOOKAHEAD_MAX + 1, &indB);
        for (int i = 0; i <= X265_LOOKAHEAD_MAX; i++) l.plannedType[i] = pt[i];
        l.indB = indB;
    }
    /* per-block arrays (frameencoder.cpp:1456,1559-1560, analysis.cpp:3681, ratecontrol.cpp), every lowres MV list the
     * lookahead published with the sentinel for the rest (search.cpp:1975-1977), and with weightp / weightb the four
     * lowres planes (weightPrediction.cpp:60-88,354-365): one asynchronous request straight into the Lowres arrays */
    pinLowres(st, f, nb);
    x265la_mirror m;
    memset(&m, 0, sizeof(m));
    m.intraCost = l.intraCost;
    if (l.qpAqOffset && l.invQscaleFactor)
    {
        m.qpAqOffset = l.qpAqOffset; m.qpCuTreeOffset = l.qpCuTreeOffset; m.invQscaleFactor = l.invQscaleFactor;
    }
    const bool isB = IS_X265_TYPE_B(l.sliceType);
    if ((st->weightPlanesP && !isB) || st->weightPlanesB)
        m.planes = l.buffer[0];             /* the 4 contiguous planes incl. margins, exactly Lowres::buffer[0..3] */
    for (int list = 0; list < 2; list++)
        for (int d = 1; d < nb && d < 18; d++)
            m.lowresMvs[list][d] = (int32_t*)l.lowresMvs[list][d];      /* MV = { int32 x, y } (common/mv.h:38-47) */
    m.d0 = -1;
    uint32_t published[2] = { 0, 0 };
    int64_t ticket = -1;
    if (x265la_frame_mirror_async(st->la, info.handle, &m, published, &ticket) != 0)
        x265_log(p, X265_LOG_ERROR, "ENABLE_CUDA lookahead: mirror failed: %s\n", x265la_last_error(st->la));
    else
        st->mirrorTicket[f] = ticket;
    for (int list = 0; list < 2; list++)
        for (int d = 0; d < nb; d++)
            if (!d || !(published[list] & (1u << d)))
                l.lowresMvs[list][d][0].x = 0x7FFF;
    (void)ncu;
    st->unreleased.push_back(info.handle);
    self.m_inputLock.acquire();
    self.m_inputQueue.remove(*f);
    self.m_inputLock.release();
    self.m_inputCount--;
    if (p->rc.rateControlMode == X265_RC_CQP)      /* Encoder::encode will not call getEstimatedPictureCost (encoder.cpp:2366) */
    {
        waitMirror(st, f);
        st->handleOf.erase(f);
        releaseHandle(st, info.handle);
    }
    /* never hold more decided frames than the slot pool was sized for */
    while (st->unreleased.size() > 6)
    {
        void* old = st->unreleased.front();
        for (std::map<Frame*, void*>::iterator h = st->handleOf.begin(); h != st->handleOf.end(); ++h)
            if (h->second == old) { st->handleOf.erase(h); break; }
        releaseHandle(st, old);
    }
    return f;
}

/* Lookahead::getEstimatedPictureCost (slicetype.cpp:1327-1439) */
void cudaLookaheadEstimatedPictureCost(Lookahead& self, Frame* cur)
{
    CudaState* st = stateOf(self);
    x265_param* p = self.m_param;
    std::map<Frame*, void*>::iterator it = st->handleOf.find(cur);
    if (it == st->handleOf.end())
    {
        x265_log(p, X265_LOG_ERROR, "ENABLE_CUDA lookahead: getEstimatedPictureCost on a frame the lookahead no longer holds\n");
        return;
    }
    void* h = it->second;
    waitMirror(st, cur);        /* the frame encoder starts reading the Lowres right after this call */
    Slice* slice = cur->m_encData->m_slice;
    const int poc = slice->m_poc;
    int d0 = 0, d1 = 0;
    if (slice->m_sliceType == P_SLICE)
        d0 = poc - slice->m_refPOCList[0][0];
    else if (slice->m_sliceType == B_SLICE)
    {
        const int l0poc = slice->m_rps.numberOfNegativePictures ? slice->m_refPOCList[0][0] : -1;
        d0 = l0poc >= 0 ? poc - l0poc : 0;
        d1 = slice->m_refPOCList[1][0] - poc;
    }
    Lowres& l = cur->m_lowres;
    l.satdCost = x265la_estimated_picture_cost_dist(st->la, h, d0, d1);
    if (p->rc.vbvBufferSize && p->rc.vbvMaxBitrate)
    {
        const int rows = x265la_vbv_rows(st->la);
        std::vector<uint32_t> satd(rows), intra(rows);
        l.lowresCostForRc = l.lowresCosts[d0][d1];
        const bool pir = p->bIntraRefresh && slice->m_sliceType == P_SLICE;
        if (x265la_vbv_row_costs(st->la, h, pir ? (int)cur->m_encData->m_pir.pirStartCol : -1, pir ? (int)cur->m_encData->m_pir.pirEndCol : -1,
                                 &satd[0], &intra[0], l.lowresCostForRc, l.intraCost) != 0)
            x265_log(p, X265_LOG_ERROR, "ENABLE_CUDA lookahead: %s\n", x265la_last_error(st->la));
        for (int i = 0; i < rows; i++)
        {
            cur->m_encData->m_rowStat[i].satdForVbv += satd[i];
            cur->m_encData->m_rowStat[i].intraSatdForVbv += intra[i];
        }
    }
    st->handleOf.erase(it);
    releaseHandle(st, h);
}

int cudaLookaheadFindSliceType(Lookahead& self, int poc)
{
    CudaState* st = stateOf(self);
    return st ? x265la_find_slice_type(st->la, poc) : X265_TYPE_AUTO;
}

} // namespace X265_NS

#endif /* ENABLE_CUDA */
