/* la_capi.cpp -- C surface of x265cu::Lookahead (see la_capi.h). */
#include "la_capi.h"
#include "lookahead.h"
#include <string.h>
#include <stdio.h>
#include <new>

using namespace x265cu;

extern "C" {

void x265la_param_default(x265la_param* q)
{
    LookaheadParam p;
    lookaheadParamDefault(&p);
    memset(q, 0, sizeof(*q));
    q->internalBitDepth = p.internalBitDepth; q->maxCUSize = p.maxCUSize;
    q->fpsNum = p.fpsNum; q->fpsDenom = p.fpsDenom;
    q->bframes = p.bframes; q->lookaheadDepth = p.lookaheadDepth; q->bFrameAdaptive = p.bFrameAdaptive;
    q->bBPyramid = p.bBPyramid; q->scenecutThreshold = p.scenecutThreshold; q->scenecutBias = 5.0;
    q->keyframeMax = p.keyframeMax; q->keyframeMin = p.keyframeMin; q->bOpenGOP = p.bOpenGOP;
    q->bEnableWeightedPred = p.bEnableWeightedPred; q->maxNumReferences = p.maxNumReferences;
    q->aqMode = p.rc.aqMode; q->aqStrength = p.rc.aqStrength; q->cuTree = p.rc.cuTree;
    q->qCompress = p.rc.qCompress; q->qgSize = p.rc.qgSize; q->rateControlMode = p.rc.rateControlMode;
    q->extraSlots = p.extraSlots; q->speculate = p.speculate; q->asyncDepth = p.asyncDepth;
    q->pendingMax = p.pendingMax; q->batchMin = p.batchMin; q->gopLookahead = p.gopLookahead; q->radl = p.radl;
    q->bFrameBias = p.bFrameBias; q->bIntraRefresh = p.bIntraRefresh; q->lookaheadSlices = p.lookaheadSlices;
    for (int i = 0; i < 2; i++) { q->hmeSearchMethod[i] = p.hmeSearchMethod[i]; q->hmeRange[i] = p.hmeRange[i]; }
}

void x265la_mvcost_table(int32_t depth, int32_t half, uint16_t* table)
{
    std::vector<uint16_t> t;
    buildMvCostTable(t, half, depth);
    memcpy(table, &t[0], t.size() * sizeof(uint16_t));
}

void* x265la_open(const x265la_param* q, char* err, int32_t errLen)
{
    LookaheadParam p;
    lookaheadParamDefault(&p);
    p.sourceWidth = q->sourceWidth; p.sourceHeight = q->sourceHeight; p.internalBitDepth = q->internalBitDepth;
    p.maxCUSize = q->maxCUSize; p.fpsNum = q->fpsNum; p.fpsDenom = q->fpsDenom;
    p.bframes = q->bframes; p.lookaheadDepth = q->lookaheadDepth; p.bFrameAdaptive = q->bFrameAdaptive;
    p.bBPyramid = q->bBPyramid; p.bFrameBias = q->bFrameBias; p.scenecutThreshold = q->scenecutThreshold;
    p.scenecutBias = q->scenecutBias / 100.0;
    p.keyframeMax = q->keyframeMax; p.keyframeMin = q->keyframeMin; p.bOpenGOP = q->bOpenGOP;
    p.bIntraRefresh = q->bIntraRefresh; p.bEnableWeightedPred = q->bEnableWeightedPred;
    p.bEnableWeightedBiPred = q->bEnableWeightedBiPred; p.lookaheadSlices = q->lookaheadSlices;
    p.maxNumReferences = q->maxNumReferences;
    p.rc.aqMode = q->aqMode; p.rc.aqStrength = q->aqStrength; p.rc.cuTree = q->cuTree; p.rc.qCompress = q->qCompress;
    p.rc.qgSize = q->qgSize; p.rc.vbvBufferSize = q->vbvBufferSize; p.rc.vbvMaxBitrate = q->vbvMaxBitrate;
    p.rc.rateControlMode = q->rateControlMode;
    p.poolWorkers = q->poolWorkers; p.device = q->device; p.extraSlots = q->extraSlots; p.speculate = q->speculate;
    p.pinHost = q->pinHost; p.asyncDepth = q->asyncDepth;
    if (q->pendingMax > 0) p.pendingMax = q->pendingMax;
    p.shardCount = q->shardCount; p.batchMin = q->batchMin; p.gopLookahead = q->gopLookahead; p.radl = q->radl;
    p.csvLogLevel = q->csvLogLevel; p.numRowsPerSlice = q->numRowsPerSlice;
    p.bEnableFades = q->bEnableFades; p.bEnableTemporalSubLayers = q->bEnableTemporalSubLayers;
    p.bHistBasedSceneCut = q->bHistBasedSceneCut;
    p.bEnableHME = q->bEnableHME && q->sourceHeight >= 540;      /* encoder.cpp:4400-4407 */
    for (int i = 0; i < 2; i++) { p.hmeSearchMethod[i] = q->hmeSearchMethod[i]; p.hmeRange[i] = q->hmeRange[i]; }
    if (p.radl && p.bOpenGOP) p.radl = 0;      /* encoder.cpp:4361-4365 */
    if (p.radl > p.bframes) p.radl = p.bframes;
    /* the adjustments Encoder::configure makes before the Lookahead sees the params
     * (encoder.cpp:3511-3516,3730-3753): cuTree needs AQ; strength 0 without cuTree disables AQ */
    if (p.rc.aqMode == 0 && p.rc.cuTree) { p.rc.aqMode = 1; p.rc.aqStrength = 0.0; }
    if (p.rc.aqStrength == 0 && p.rc.cuTree == 0) p.rc.aqMode = 0;
    /* without AQ and VBV the quantisation group is the CTU (encoder.cpp:4108-4123) */
    if (!(p.rc.aqMode || (p.rc.vbvBufferSize > 0 && p.rc.vbvMaxBitrate > 0))) p.rc.qgSize = p.maxCUSize;
    else if (p.rc.qgSize > p.maxCUSize) p.rc.qgSize = p.maxCUSize;
    /* temporal layers (encoder.cpp:3914-3943): 1 is not a mode, more than 5 become 5, 3..5 fix the mini-GOP and turn b-adapt off */
    if (p.bEnableTemporalSubLayers > 2 && !p.bframes) p.bEnableTemporalSubLayers = 0;
    if (p.bEnableTemporalSubLayers == 1) p.bEnableTemporalSubLayers = 0;
    if (p.bEnableTemporalSubLayers > 5) p.bEnableTemporalSubLayers = 5;
    if (p.bEnableTemporalSubLayers > 2)
    {
        static const int tlBframes[6] = { 0, 0, 0, 3, 7, 15 };
        p.bframes = tlBframes[p.bEnableTemporalSubLayers];
        p.bFrameAdaptive = 0;
    }
    if (!p.bframes) p.bBPyramid = 0;
    if (p.bIntraRefresh)        /* encoder.cpp:3761-3779 */
    {
        if (p.maxNumReferences > 1) p.maxNumReferences = 1;
        p.bBPyramid = 0;
        p.bOpenGOP = 0;
    }
    Lookahead* la = new (std::nothrow) Lookahead(p);
    if (!la) { if (err && errLen) snprintf(err, errLen, "out of memory"); return NULL; }
    if (!la->create())
    {
        if (err && errLen) snprintf(err, errLen, "%s", la->lastError());
        delete la;
        return NULL;
    }
    return la;
}

void x265la_close(void* la) { delete (Lookahead*)la; }

int x265la_get_geometry(void* la, x265cu_geometry* g) { *g = ((Lookahead*)la)->geometry(); return 0; }
const char* x265la_last_error(void* la) { return ((Lookahead*)la)->lastError(); }
x265cu_ctx* x265la_engine(void* la) { return ((Lookahead*)la)->engine(); }
int x265la_shard_config(void* la, int32_t rank, int32_t nranks, x265cu_exchange_fn fn, void* user)
{
    ((Lookahead*)la)->setShardRank(rank);
    return x265cu_shard_config(((Lookahead*)la)->engine(), rank, nranks, fn, user);
}

void* x265la_add_picture(void* la, const void* y, const void* u, const void* v, int32_t strideY, int32_t strideC,
                         int64_t pts, int32_t sliceType, int32_t sliceTypeReq)
{
    return ((Lookahead*)la)->addPicture(y, u, v, strideY, strideC, pts, sliceType, sliceTypeReq);
}

void x265la_flush(void* la) { ((Lookahead*)la)->flush(); }

int x265la_get_decided(void* lav, x265la_frame_info* out)
{
    Lookahead* la = (Lookahead*)lav;
    Frame* f = la->getDecidedPicture();
    if (!la->ok()) return -1;
    if (!f) return 0;
    const Lowres& l = f->m_lowres;
    out->poc = f->m_poc; out->sliceType = l.sliceType; out->bScenecut = l.bScenecut; out->bKeyframe = l.bKeyframe;
    out->bLastMiniGopBFrame = l.bLastMiniGopBFrame; out->leadingBframes = l.leadingBframes;
    out->pts = f->m_pts; out->reorderedPts = f->m_reorderedPts; out->satdCost = l.satdCost;
    out->gopOffset = f->m_gopOffset; out->gopId = f->m_gopId; out->tempLayer = f->m_tempLayer; out->gopIdWritten = f->m_gopIdSet;
    out->handle = f;
    return 1;
}

int64_t x265la_estimated_picture_cost(void* la, void* frame, void* ref0, void* ref1)
{
    ((Lookahead*)la)->getEstimatedPictureCost((Frame*)frame, (Frame*)ref0, (Frame*)ref1);
    return ((Frame*)frame)->m_lowres.satdCost;
}

int64_t x265la_estimated_picture_cost_dist(void* la, void* frame, int32_t d0, int32_t d1)
{
    ((Lookahead*)la)->getEstimatedPictureCost((Frame*)frame, d0, d1);
    return ((Frame*)frame)->m_lowres.satdCost;
}

int x265la_find_slice_type(void* la, int32_t poc) { return ((Lookahead*)la)->findSliceType(poc); }
double x265la_frame_ip_cost_ratio(void* /*la*/, void* frame) { return ((Frame*)frame)->m_lowres.ipCostRatio; }

void x265la_release(void* la, void* frame) { ((Lookahead*)la)->releaseFrame((Frame*)frame); }

int x265la_vbv_rows(void* la) { return ((Lookahead*)la)->vbvRows(); }

int x265la_vbv_row_costs(void* la, void* frame, int32_t pirStartCol, int32_t pirEndCol, uint32_t* satdForVbv,
                         uint32_t* intraSatdForVbv, uint16_t* lowresCostForRc, int32_t* intraCostScaled)
{
    return ((Lookahead*)la)->getVbvRowCosts((Frame*)frame, pirStartCol, pirEndCol, satdForVbv, intraSatdForVbv,
                                            lowresCostForRc, intraCostScaled) ? 0 : -1;
}

int x265la_frame_planned(void* /*la*/, void* frame, int64_t* plannedSatd, int32_t* plannedType, int32_t n, int32_t* indB)
{
    const Lowres& l = ((Frame*)frame)->m_lowres;
    for (int i = 0; i < n && i <= LOOKAHEAD_MAX; i++)
    {
        if (plannedSatd) plannedSatd[i] = l.plannedSatd[i];
        if (plannedType) plannedType[i] = l.plannedType[i];
    }
    if (indB) *indB = l.indB;
    return 0;
}

int x265la_frame_scalars(void* lav, void* frame, int64_t* costEst, int64_t* costEstAq, int32_t* intraMbs,
                         int32_t* rowSatdsValid, uint64_t* wp_ssd, uint64_t* wp_sum, double* wdelta)
{
    Lookahead* la = (Lookahead*)lav;
    const Lowres& l = ((Frame*)frame)->m_lowres;
    const int nb = la->geometry().nb;
    for (int i = 0; i < nb; i++)
    {
        for (int j = 0; j < nb; j++)
        {
            if (costEst) costEst[i * nb + j] = l.costEst[i][j];
            if (costEstAq) costEstAq[i * nb + j] = l.costEstAq[i][j];
            if (rowSatdsValid) rowSatdsValid[i * nb + j] = l.rowSatdsValid[i][j];
        }
        if (intraMbs) intraMbs[i] = l.intraMbs[i];
        if (wdelta) wdelta[i] = l.weightedCostDelta[i];
    }
    for (int k = 0; k < 3; k++) { if (wp_ssd) wp_ssd[k] = l.wp_ssd[k]; if (wp_sum) wp_sum[k] = l.wp_sum[k]; }
    return 0;
}

int x265la_frame_hist(void*, void* frame, int32_t* variance, int32_t* intensity, uint64_t* checksum)
{
    const Lowres& l = ((Frame*)frame)->m_lowres;
    if (!l.hist) return -1;
    uint64_t ck = 0;
    for (int i = 0; i < 3; i++) { variance[i] = l.hist->pic_avg_variance[i]; intensity[i] = l.hist->avg_intensity[i]; }
    for (int wi = 0; wi < 4; wi++)
        for (int hi = 0; hi < 4; hi++)
            for (int pl = 0; pl < 3; pl++)
            {
                ck = ck * 1000003u + l.hist->avg_intensity_seg[wi][hi][pl];
                for (int b = 0; b < 256; b++) ck = ck * 1000003u + l.hist->histogram[wi][hi][pl][b];
            }
    *checksum = ck;
    return 0;
}

int x265la_frame_fade(void*, void* frame, int32_t* bIsFadeEnd, double* frameVariance)
{
    const Lowres& l = ((Frame*)frame)->m_lowres;
    if (bIsFadeEnd) *bIsFadeEnd = l.bIsFadeEnd;
    if (frameVariance) *frameVariance = l.frameVariance;
    return 0;
}

int x265la_frame_mvs(void* la, void* frame, int32_t list, int32_t dist, int32_t* mvXY, int32_t* mvCosts)
{ return ((Lookahead*)la)->fetchMvs((Frame*)frame, list, dist, mvXY, mvCosts) ? 1 : 0; }

int x265la_frame_hme_mvs(void* la, void* frame, int32_t list, int32_t dist, int32_t* mvXY, int32_t* mvCosts)
{ return ((Lookahead*)la)->fetchHmeMvs((Frame*)frame, list, dist, mvXY, mvCosts) ? 1 : 0; }

int x265la_frame_costs(void* la, void* frame, int32_t d0, int32_t d1, uint16_t* lowresCosts, int32_t* rowSatds)
{ return ((Lookahead*)la)->fetchCosts((Frame*)frame, d0, d1, lowresCosts, rowSatds) ? 1 : 0; }

int x265la_frame_mirror_async(void* lav, void* frame, const x265la_mirror* m, uint32_t published[2], int64_t* ticket)
{
    Lookahead* la = (Lookahead*)lav;
    Frame* f = (Frame*)frame;
    const Lowres& l = f->m_lowres;
    const int nb = la->geometry().nb;
    x265cu_mirror_request q;
    memset(&q, 0, sizeof(q));
    q.intra_cost = m->intraCost; q.qp_aq_offset = m->qpAqOffset; q.qp_cutree_offset = m->qpCuTreeOffset;
    q.inv_qscale_factor = m->invQscaleFactor; q.planes = m->planes;
    if (published) published[0] = published[1] = 0;
    for (int list = 0; list < 2; list++)
        for (int d = 1; d < nb && d < 18; d++)
            if (m->lowresMvs[list][d] && l.mvStore[list][d] >= 0 && q.n_mv < X265CU_MIRROR_MAX_MV)
            {
                q.mv_store[q.n_mv] = l.mvStore[list][d]; q.mv_dst[q.n_mv] = m->lowresMvs[list][d]; q.n_mv++;
                if (published) published[list] |= 1u << d;
            }
    q.cost_store = -1;
    if (m->d0 >= 0 && m->d0 < nb && m->d1 >= 0 && m->d1 < nb && l.costStore[m->d0][m->d1] >= 0)
    {
        q.cost_store = l.costStore[m->d0][m->d1]; q.lowres_costs = m->lowresCosts; q.row_satds = m->rowSatds;
    }
    return la->mirror(f, &q, ticket) ? 0 : -1;
}

int x265la_mirror_wait(void* la, int64_t ticket) { return x265cu_mirror_wait(((Lookahead*)la)->engine(), ticket); }
int x265la_pin(void* la, void* ptr, uint64_t bytes) { return x265cu_pin_host(((Lookahead*)la)->engine(), ptr, bytes); }
int x265la_unpin(void* la, void* ptr) { return x265cu_unpin_host(((Lookahead*)la)->engine(), ptr); }

int x265la_frame_weights(void* lav, void* frame, int32_t* state, int32_t* scale, int32_t* denom, int32_t* offset)
{
    const Lowres& l = ((Frame*)frame)->m_lowres;
    const int nb = ((Lookahead*)lav)->geometry().nb;
    for (int i = 0; i < nb; i++)
    {
        state[i] = l.weightState[i]; scale[i] = l.wScale[i]; denom[i] = l.wDenom[i]; offset[i] = l.wOffset[i];
    }
    return 0;
}

int x265la_get_timers(void* lav, double* t, int32_t reset)
{
    Lookahead* la = (Lookahead*)lav;
    for (int i = 0; i < 10; i++) t[i] = la->m_timers[i];
    if (reset) memset(la->m_timers, 0, sizeof(la->m_timers));
    return 0;
}

int x265la_frame_fetch(void* la, void* frame, const x265cu_frame_out* out)
{ return ((Lookahead*)la)->fetchFrame((Frame*)frame, out) ? 0 : -1; }

} // extern "C"
