/* ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the UNMODIFIED reference lookahead (DJATOM/x265-aMod 3.6+1) compiled from
 * /root/reference/source by oracle/Makefile.ref, and exposes every Lowres field the
 * hot path produces through a tiny C ABI so tests can pin the CUDA path (and the C
 * restatement in oracle/la_oracle.c) against the real thing.
 *
 * How: a real Encoder is opened (so x265_param is configured exactly as the encoder
 * would, encoder/encoder.cpp:3608-...), then frames are pushed straight into
 * Encoder::m_lookahead with Lookahead::addPicture (encoder/slicetype.cpp:1200) and
 * drained with Lookahead::getDecidedPicture (:1289) -- Encoder::encode is never
 * called.  Each drained frame's Lowres state is snapshotted.
 *
 * This file contains no reference source text; it only calls the reference's classes. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <vector>
#include <string>
#include <new>
#include <time.h>

/* the harness needs Lookahead::frameCostRecalculate / m_tld (protected) and the
 * BitCost tables (protected static) */
#define protected public
#define private public
#include "common.h"
#include "x265.h"
#include "param.h"
#include "primitives.h"
#include "frame.h"
#include "framedata.h"
#include "picyuv.h"
#include "lowres.h"
#include "bitcost.h"
#include "threadpool.h"
#include "slicetype.h"
#include "encoder.h"
#undef protected
#undef private

using namespace X265_NS;

extern "C" {

struct RefLaConfig
{
    int32_t width, height;          /* picture size as handed to the encoder */
    int32_t fpsNum, fpsDenom;
    int32_t bframes, lookaheadDepth, bFrameAdaptive, bBPyramid;
    int32_t scenecutThreshold;      /* 0 disables */
    int32_t keyframeMax, keyframeMin; /* keyframeMin 0 = auto */
    int32_t bOpenGOP;
    int32_t aqMode;
    double  aqStrength;
    int32_t cuTree;
    double  qCompress;
    int32_t weightp, weightb;
    int32_t poolThreads;            /* 0 = no pool (everything on the caller) */
    int32_t lookaheadSlices;
    int32_t qgSize;                 /* 8/16/32/64 */
    int32_t bFrameBias;
    double  scenecutBias;           /* percent, as --scenecut-bias */
    int32_t vbvBufferSize, vbvMaxBitrate, bitrate; /* 0 = CRF */
    int32_t dumpPlanes;             /* keep the 4 lowres planes of every frame */
    int32_t bIntraRefresh;
    int32_t gopLookahead;           /* --gop-lookahead */
    int32_t radl;                   /* --radl */
    int32_t keepFrames;             /* drained frames kept alive (0 = just what the lookahead itself still references);
                                       ref_la_estimate needs a frame and its references alive */
    int32_t fades;                  /* --fades (x265_param::bEnableFades) */
    int32_t temporalLayers;         /* --temporal-layers (x265_param::bEnableTemporalSubLayers) */
    int32_t histScenecut;           /* --hist-scenecut (x265_param::bHistBasedSceneCut) */
    int32_t csp400;                 /* 1 = 4:0:0 (x265_param::internalCsp = X265_CSP_I400): pictures carry no chroma */
    int32_t hme;                    /* --hme (x265_param::bEnableHME; the encoder turns it off below 540 lines) */
    int32_t hmeSearch0, hmeSearch1; /* --hme-search levels 0 and 1 (level 2 is the main encoder's) */
    int32_t hmeRange0, hmeRange1;   /* --hme-range levels 0 and 1 */
};

struct RefLaFrame
{
    int32_t poc, sliceType, bScenecut, bKeyframe, bLastMiniGopBFrame, leadingBframes;
    int32_t bw, bh, nb;             /* 8x8 grid, nb = bframes+2 */
    int32_t stride, planeLines;     /* lowres plane geometry (pixels) */
    int64_t satdCost;
    const int64_t*  costEst;        /* nb*nb */
    const int64_t*  costEstAq;      /* nb*nb */
    const int32_t*  intraMbs;       /* nb */
    const int32_t*  rowSatds;       /* nb*nb*bh (row 0 == -1 => never computed) */
    const uint16_t* lowresCosts;    /* nb*nb*ncu */
    const int32_t*  mvs;            /* 2*nb*ncu*2 (x,y) */
    const int32_t*  mvCosts;        /* 2*nb*ncu */
    const int32_t*  intraCost;      /* ncu */
    const uint8_t*  intraMode;      /* ncu */
    const double*   qpAqOffset;     /* ncu */
    const double*   qpCuTreeOffset; /* ncu */
    const int32_t*  invQscaleFactor;/* ncu */
    const uint16_t* propagateCost;  /* ncu */
    uint64_t wp_ssd[3], wp_sum[3];
    const double*   weightedCostDelta; /* nb */
    const void*     planes;         /* 4*stride*planeLines pixels or NULL */
    int32_t ncuFull;                /* entries of qpAqOffset / qpCuTreeOffset / invQscaleFactor (4*ncu with qg-size 8) */
    int32_t indB;                   /* Lowres::indB */
    const int64_t*  plannedSatd;    /* X265_LOOKAHEAD_MAX + 1 */
    const int32_t*  plannedType;    /* X265_LOOKAHEAD_MAX + 1 */
    /* filled by ref_la_estimate (Lookahead::getEstimatedPictureCost on this frame), else estimated == 0 */
    int32_t estimated, vbvRows;
    int64_t estSatdCost;            /* Lowres::satdCost afterwards */
    const uint32_t* satdForVbv;     /* vbvRows: FrameData::m_rowStat[].satdForVbv */
    const uint32_t* intraSatdForVbv;
    const uint16_t* lowresCostForRc;/* ncu, after the in-place scaling of slicetype.cpp:1411-1428 */
    const int32_t*  intraCostForRc; /* ncu, Lowres::intraCost after the same */
    const int32_t*  estRowSatds;    /* bh: rowSatds of the coded estimate after frameCostRecalculate */
    int32_t bIsFadeEnd, pad0;       /* Lowres::bIsFadeEnd (--fades) */
    double  frameVariance;          /* Lowres::frameVariance (--fades) */
    /* --hist-scenecut: Lowres::picAvgVariance{,Cb,Cr}, averageIntensity[3], and a checksum of picHistogram + averageIntensityPerSegment */
    int32_t histVar[3], histAvg[3];
    uint64_t histCheck;
    /* --hme: Lowres::lowerResMvs / lowerResMvCosts (the level-0 searches), else NULL */
    int32_t bw4, bh4;
    const int32_t*  lowerMvs;       /* 2*nb*ncu4*2 (x,y) */
    const int32_t*  lowerMvCosts;   /* 2*nb*ncu4 */
    /* --temporal-layers 3..5: Frame::m_gopOffset / m_gopId / m_tempLayer as the decision left them */
    int32_t gopOffset, gopId, tempLayer, pad1;
};

} // extern "C"

namespace {

struct FrameSnap
{
    RefLaFrame h;
    std::vector<int64_t> costEst, costEstAq;
    std::vector<int32_t> intraMbs, rowSatds, mvs, mvCosts, intraCost, invQ;
    std::vector<uint16_t> lowresCosts, propagate;
    std::vector<uint8_t> intraMode;
    std::vector<double> qpAq, qpCuTree, wdelta;
    std::vector<pixel> planes;
    std::vector<int64_t> plannedSatd;
    std::vector<int32_t> plannedType, intraCostForRc, estRowSatds;
    std::vector<uint32_t> satdForVbv, intraSatdForVbv;
    std::vector<uint16_t> lowresCostForRc;
    std::vector<int32_t> lowerMvs, lowerMvCosts;
    Frame* frame;                    /* NULL once the frame was destroyed */
};

struct Handle
{
    RefLaConfig cfg;
    x265_param* userParam;
    Encoder*    enc;
    Lookahead*  la;
    int         pocNext;
    bool        flushed;
    std::vector<FrameSnap*> out;
    std::vector<Frame*> retired;     /* frames drained from the lookahead; kept alive because
                                        later decisions still reference m_lastNonB */
    double      secondsInLookahead;
    double      secondsInSnapshot;   /* harness overhead inside put / flush, for callers that time those calls */
};

void snapshot(Handle* h, Frame* f)
{
    Lowres& l = f->m_lowres;
    FrameSnap* s = new FrameSnap;
    const int nb = h->enc->m_param->bframes + 2;
    const int bw = l.maxBlocksInRow, bh = l.maxBlocksInCol, ncu = bw * bh;
    memset(&s->h, 0, sizeof(s->h));
    s->h.poc = f->m_poc; s->h.sliceType = l.sliceType; s->h.bScenecut = l.bScenecut;
    s->h.bKeyframe = l.bKeyframe; s->h.bLastMiniGopBFrame = l.bLastMiniGopBFrame;
    s->h.leadingBframes = l.leadingBframes;
    s->h.bIsFadeEnd = l.bIsFadeEnd; s->h.frameVariance = h->enc->m_param->bEnableFades ? l.frameVariance : 0;
    if (h->enc->m_param->bHistBasedSceneCut && l.picHistogram)
    {
        s->h.histVar[0] = l.picAvgVariance; s->h.histVar[1] = l.picAvgVarianceCb; s->h.histVar[2] = l.picAvgVarianceCr;
        uint64_t ck = 0;
        for (int i = 0; i < 3; i++) s->h.histAvg[i] = l.averageIntensity[i];
        for (int wi = 0; wi < 4; wi++)
            for (int hi = 0; hi < 4; hi++)
                for (int pl = 0; pl < 3; pl++)
                {
                    ck = ck * 1000003u + (uint8_t)l.averageIntensityPerSegment[wi][hi][pl];
                    for (int b = 0; b < 256; b++) ck = ck * 1000003u + l.picHistogram[wi][hi][pl][b];
                }
        s->h.histCheck = ck;
    }
    s->h.gopOffset = f->m_gopOffset; s->h.gopId = f->m_gopId; s->h.tempLayer = f->m_tempLayer;
    s->h.bw = bw; s->h.bh = bh; s->h.nb = nb;
    s->h.stride = (int)l.lumaStride;
    s->h.planeLines = (int)((l.buffer[1] - l.buffer[0]) / l.lumaStride);
    s->h.satdCost = l.satdCost;
    for (int i = 0; i < nb; i++)
        for (int j = 0; j < nb; j++)
        {
            s->costEst.push_back(l.costEst[i][j]);
            s->costEstAq.push_back(l.costEstAq[i][j]);
            s->rowSatds.insert(s->rowSatds.end(), l.rowSatds[i][j], l.rowSatds[i][j] + bh);
            s->lowresCosts.insert(s->lowresCosts.end(), l.lowresCosts[i][j], l.lowresCosts[i][j] + ncu);
        }
    for (int i = 0; i < nb; i++)
    {
        s->intraMbs.push_back(l.intraMbs[i]);
        s->wdelta.push_back(l.weightedCostDelta[i]);
    }
    for (int list = 0; list < 2; list++)
        for (int i = 0; i < nb; i++)
        {
            for (int c = 0; c < ncu; c++)
            {
                s->mvs.push_back(l.lowresMvs[list][i][c].x);
                s->mvs.push_back(l.lowresMvs[list][i][c].y);
            }
            s->mvCosts.insert(s->mvCosts.end(), l.lowresMvCosts[list][i], l.lowresMvCosts[list][i] + ncu);
        }
    if (h->enc->m_param->bEnableHME)
    {
        const int bw4 = ((l.width / 2) + X265_LOWRES_CU_SIZE - 1) >> X265_LOWRES_CU_BITS;     /* lowres.cpp:205-207 */
        const int bh4 = ((l.lines / 2) + X265_LOWRES_CU_SIZE - 1) >> X265_LOWRES_CU_BITS;
        s->h.bw4 = bw4; s->h.bh4 = bh4;
        for (int list = 0; list < 2; list++)
            for (int i = 0; i < nb; i++)
            {
                for (int c = 0; c < bw4 * bh4; c++)
                {
                    s->lowerMvs.push_back(l.lowerResMvs[list][i][c].x);
                    s->lowerMvs.push_back(l.lowerResMvs[list][i][c].y);
                }
                s->lowerMvCosts.insert(s->lowerMvCosts.end(), l.lowerResMvCosts[list][i], l.lowerResMvCosts[list][i] + bw4 * bh4);
            }
        s->h.lowerMvs = s->lowerMvs.data(); s->h.lowerMvCosts = s->lowerMvCosts.data();
    }
    s->intraCost.assign(l.intraCost, l.intraCost + ncu);
    s->intraMode.assign(l.intraMode, l.intraMode + ncu);
    const int ncuFull = h->enc->m_param->rc.qgSize == 8 ? 4 * ncu : ncu;      /* lowres.cpp:89 */
    s->h.ncuFull = ncuFull;
    if (l.qpAqOffset)
    {
        s->qpAq.assign(l.qpAqOffset, l.qpAqOffset + ncuFull);
        s->qpCuTree.assign(l.qpCuTreeOffset, l.qpCuTreeOffset + ncuFull);
        s->invQ.assign(l.invQscaleFactor, l.invQscaleFactor + ncuFull);
    }
    else
    {
        s->qpAq.assign(ncuFull, 0.0); s->qpCuTree.assign(ncuFull, 0.0); s->invQ.assign(ncuFull, 256);
    }
    s->h.indB = l.indB;
    s->plannedSatd.assign(l.plannedSatd, l.plannedSatd + X265_LOOKAHEAD_MAX + 1);
    s->plannedType.assign(l.plannedType, l.plannedType + X265_LOOKAHEAD_MAX + 1);
    s->h.plannedSatd = s->plannedSatd.data(); s->h.plannedType = s->plannedType.data();
    s->frame = f;
    s->propagate.assign(l.propagateCost, l.propagateCost + ncu);
    for (int i = 0; i < 3; i++) { s->h.wp_ssd[i] = l.wp_ssd[i]; s->h.wp_sum[i] = l.wp_sum[i]; }
    if (h->cfg.dumpPlanes)
        s->planes.assign(l.buffer[0], l.buffer[0] + 4 * (size_t)s->h.stride * s->h.planeLines);

    s->h.costEst = s->costEst.data(); s->h.costEstAq = s->costEstAq.data();
    s->h.intraMbs = s->intraMbs.data(); s->h.rowSatds = s->rowSatds.data();
    s->h.lowresCosts = s->lowresCosts.data(); s->h.mvs = s->mvs.data(); s->h.mvCosts = s->mvCosts.data();
    s->h.intraCost = s->intraCost.data(); s->h.intraMode = s->intraMode.data();
    s->h.qpAqOffset = s->qpAq.data(); s->h.qpCuTreeOffset = s->qpCuTree.data();
    s->h.invQscaleFactor = s->invQ.data(); s->h.propagateCost = s->propagate.data();
    s->h.weightedCostDelta = s->wdelta.data();
    s->h.planes = s->planes.empty() ? NULL : (const void*)s->planes.data();
    h->out.push_back(s);
}

double nowSec()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

void drain(Handle* h, bool snap)
{
    for (;;)
    {
        double t0 = nowSec();
        Frame* f = h->la->getDecidedPicture();
        h->secondsInLookahead += nowSec() - t0;
        if (!f)
            break;
        if (snap)
        {
            double s0 = nowSec();
            snapshot(h, f);
            h->secondsInSnapshot += nowSec() - s0;
        }
        else
        {
            FrameSnap* s = new FrameSnap;
            memset(&s->h, 0, sizeof(s->h));
            s->frame = f;
            s->h.poc = f->m_poc; s->h.sliceType = f->m_lowres.sliceType;
            s->h.bScenecut = f->m_lowres.bScenecut; s->h.bKeyframe = f->m_lowres.bKeyframe;
            h->out.push_back(s);
        }
        h->retired.push_back(f);
        /* keep a bounded tail alive: m_lastNonB and cuTree only ever look at the most
         * recent non-B, so anything older than 2*(bframes+2) drained frames is dead */
        size_t keep = 2 * (size_t)(h->enc->m_param->bframes + 2) + 2;
        if (h->cfg.keepFrames > (int)keep) keep = h->cfg.keepFrames;
        while (h->retired.size() > keep)
        {
            Frame* old = h->retired.front();
            h->retired.erase(h->retired.begin());
            for (size_t k = 0; k < h->out.size(); k++)
                if (h->out[k]->frame == old) h->out[k]->frame = NULL;
            old->destroy();
            delete old;
        }
    }
}

} // namespace

extern "C" {

int ref_la_depth(void) { return X265_DEPTH; }

void* ref_la_open(const RefLaConfig* c)
{
    x265_param* p = PARAM_NS::x265_param_alloc();
    if (!p) return NULL;
    PARAM_NS::x265_param_default_preset(p, "medium", NULL);
    p->sourceWidth = c->width; p->sourceHeight = c->height;
    p->fpsNum = c->fpsNum; p->fpsDenom = c->fpsDenom;
    p->internalCsp = c->csp400 ? X265_CSP_I400 : X265_CSP_I420;
    p->internalBitDepth = X265_DEPTH;
    p->sourceBitDepth = X265_DEPTH;
    p->logLevel = getenv("REF_LA_LOG") ? atoi(getenv("REF_LA_LOG")) : X265_LOG_NONE;
    p->bframes = c->bframes;
    p->lookaheadDepth = c->lookaheadDepth;
    p->bFrameAdaptive = c->bFrameAdaptive;
    p->bBPyramid = c->bBPyramid;
    p->scenecutThreshold = c->scenecutThreshold;
    p->scenecutBias = c->scenecutBias; /* percent; Encoder::configure divides by 100 (encoder.cpp:3948) */
    p->keyframeMax = c->keyframeMax;
    p->keyframeMin = c->keyframeMin;
    p->bOpenGOP = c->bOpenGOP;
    p->rc.aqMode = c->aqMode;
    p->rc.aqStrength = c->aqStrength;
    p->rc.cuTree = c->cuTree;
    p->rc.qCompress = c->qCompress;
    p->rc.qgSize = c->qgSize;
    p->bEnableWeightedPred = c->weightp;
    p->bEnableWeightedBiPred = c->weightb;
    p->lookaheadSlices = c->lookaheadSlices;
    p->bFrameBias = c->bFrameBias;
    p->bIntraRefresh = c->bIntraRefresh;
    p->gopLookahead = c->gopLookahead;
    p->radl = c->radl;
    p->bEnableFades = c->fades;
    p->bEnableTemporalSubLayers = c->temporalLayers;
    p->bHistBasedSceneCut = c->histScenecut;
    if (c->hme)
    {
        p->bEnableHME = 1;
        p->hmeSearchMethod[0] = c->hmeSearch0; p->hmeSearchMethod[1] = c->hmeSearch1;
        p->hmeRange[0] = c->hmeRange0; p->hmeRange[1] = c->hmeRange1;
    }
    if (c->vbvBufferSize)
    {
        p->rc.rateControlMode = X265_RC_ABR;
        p->rc.bitrate = c->bitrate;
        p->rc.vbvBufferSize = c->vbvBufferSize;
        p->rc.vbvMaxBitrate = c->vbvMaxBitrate;
    }
    static char poolStr[32];
    char* ps = (char*)malloc(32);
    if (c->poolThreads > 0) snprintf(ps, 32, "%d", c->poolThreads);
    else snprintf(ps, 32, "none");
    p->numaPools = ps;
    p->frameNumThreads = 1;
    (void)poolStr;

    x265_encoder* e = x265_encoder_open(p);
    if (!e) { PARAM_NS::x265_param_free(p); return NULL; }
    Handle* h = new Handle;
    h->cfg = *c; h->userParam = p;
    h->enc = static_cast<Encoder*>(e);
    h->la = h->enc->m_lookahead;
    h->pocNext = 0; h->flushed = false; h->secondsInLookahead = 0; h->secondsInSnapshot = 0;
    return h;
}

/* effective (encoder-configured) parameters the other side must mirror */
void ref_la_effective(void* hv, int32_t* out /* [16] */)
{
    Handle* h = (Handle*)hv;
    x265_param* p = h->enc->m_param;
    out[0] = p->sourceWidth; out[1] = p->sourceHeight; out[2] = p->keyframeMin; out[3] = p->keyframeMax;
    out[4] = p->bframes; out[5] = p->lookaheadDepth; out[6] = p->bFrameAdaptive; out[7] = p->bBPyramid;
    out[8] = p->rc.aqMode; out[9] = p->rc.cuTree; out[10] = p->rc.qgSize; out[11] = p->lookaheadSlices;
    out[12] = h->la->m_pool ? h->la->m_pool->m_numWorkers : 0;
    out[13] = p->bEnableWeightedPred; out[14] = p->bEnableWeightedBiPred; out[15] = p->scenecutThreshold;
}

/* push one 4:2:0 picture (pixel = uint8_t or uint16_t per this library's depth);
 * strides in pixels.  Returns total frames decided so far. */
int ref_la_put_typed(void* hv, const void* y, const void* u, const void* v, int strideY, int strideC, int snap, int sliceType,
                     int sliceTypeReq);

int ref_la_put(void* hv, const void* y, const void* u, const void* v, int strideY, int strideC, int snap)
{
    return ref_la_put_typed(hv, y, u, v, strideY, strideC, snap, X265_TYPE_AUTO, X265_TYPE_AUTO);
}

/* The two ways a slice type reaches the lookahead, exactly as Encoder::encode does it (encoder.cpp:1713-1714, 1863):
 * sliceTypeReq = x265_picture::sliceType forced by the application -> Frame::m_lowres.sliceTypeReq;
 * sliceType    = the first-pass type of a 2-pass encode            -> the argument of Lookahead::addPicture */
int ref_la_put_typed(void* hv, const void* y, const void* u, const void* v, int strideY, int strideC, int snap, int sliceType,
                     int sliceTypeReq)
{
    Handle* h = (Handle*)hv;
    x265_param* p = h->enc->m_param;
    Frame* f = new Frame;
    if (!f->create(p, NULL)) return -1;
    x265_picture pic;
    x265_picture_init(p, &pic);
    pic.bitDepth = X265_DEPTH;
    pic.colorSpace = h->cfg.csp400 ? X265_CSP_I400 : X265_CSP_I420;
    pic.planes[0] = (void*)y; pic.planes[1] = h->cfg.csp400 ? NULL : (void*)u; pic.planes[2] = h->cfg.csp400 ? NULL : (void*)v;
    pic.stride[0] = strideY * (int)sizeof(pixel);
    pic.stride[1] = pic.stride[2] = strideC * (int)sizeof(pixel);
    /* conformance-window padding exactly as Encoder::encode passes it (encoder.cpp:1645) */
    f->m_fencPic->copyFromPicture(pic, *p, h->enc->m_sps.conformanceWindow.rightOffset,
                                  h->enc->m_sps.conformanceWindow.bottomOffset);
    f->m_poc = h->pocNext;
    f->m_pts = h->pocNext;
    h->pocNext++;
    f->m_lowres.sliceTypeReq = sliceTypeReq;
    f->m_lowres.bScenecut = false;
    f->m_lowres.satdCost = (int64_t)-1;
    f->m_lowresInit = false;
    double t0 = nowSec();
    h->la->addPicture(*f, sliceType);
    h->secondsInLookahead += nowSec() - t0;
    drain(h, snap != 0);
    return (int)h->out.size();
}

int ref_la_flush(void* hv, int snap)
{
    Handle* h = (Handle*)hv;
    h->la->flush();
    h->flushed = true;
    /* getDecidedPicture returns NULL only once both queues are empty */
    for (;;)
    {
        size_t before = h->out.size();
        drain(h, snap != 0);
        if (h->out.size() == before)
            break;
    }
    return (int)h->out.size();
}

/* Lookahead::getEstimatedPictureCost (slicetype.cpp:1327-1439) on decided frame `idx` (output order) with the list-0 /
 * list-1 references `idxRef0` / `idxRef1` (-1 = none), the way Encoder::encode calls it for the frame about to be coded
 * (encoder.cpp:2367).  The reference reads the references from the slice header, so a minimal FrameData / Slice is put
 * in front of it.  The frame's satdCost, the per-CTU-row VBV sums and the arrays the call rescales in place are added
 * to the frame's snapshot.  Returns 0, or -1 when one of the frames is no longer alive. */
int ref_la_estimate(void* hv, int idx, int idxRef0, int idxRef1, int pirStartCol, int pirEndCol)
{
    Handle* h = (Handle*)hv;
    const int n = (int)h->out.size();
    if (idx < 0 || idx >= n || idxRef0 >= n || idxRef1 >= n) return -1;
    FrameSnap* s = h->out[idx];
    Frame* cur = s->frame;
    Frame* r0 = idxRef0 >= 0 ? h->out[idxRef0]->frame : NULL;
    Frame* r1 = idxRef1 >= 0 ? h->out[idxRef1]->frame : NULL;
    if (!cur || (idxRef0 >= 0 && !r0) || (idxRef1 >= 0 && !r1)) return -1;
    x265_param* p = h->enc->m_param;
    const int rows = (p->sourceHeight + p->maxCUSize - 1) / p->maxCUSize;
    FrameData fd;
    Slice slice;
    std::vector<FrameData::RCStatRow> rowStat(rows);
    memset(rowStat.data(), 0, rows * sizeof(FrameData::RCStatRow));
    fd.m_slice = &slice; fd.m_rowStat = rowStat.data();
    fd.m_pir.pirStartCol = pirStartCol; fd.m_pir.pirEndCol = pirEndCol;
    const int t = cur->m_lowres.sliceType;
    slice.m_sliceType = IS_X265_TYPE_I(t) ? I_SLICE : (t == X265_TYPE_P ? P_SLICE : B_SLICE);
    slice.m_poc = cur->m_poc;
    slice.m_rps.numberOfNegativePictures = r0 ? 1 : 0;
    if (r0) { slice.m_refPOCList[0][0] = r0->m_poc; slice.m_refFrameList[0][0] = r0; }
    if (r1) { slice.m_refPOCList[1][0] = r1->m_poc; slice.m_refFrameList[1][0] = r1; }
    FrameData* saved = cur->m_encData;
    cur->m_encData = &fd;
    h->la->getEstimatedPictureCost(cur);
    cur->m_encData = saved;
    fd.m_slice = NULL; fd.m_rowStat = NULL;
    Lowres& l = cur->m_lowres;
    const int ncu = l.maxBlocksInRow * l.maxBlocksInCol;
    s->h.estimated = 1; s->h.vbvRows = rows; s->h.estSatdCost = l.satdCost;
    s->satdForVbv.resize(rows); s->intraSatdForVbv.resize(rows);
    for (int i = 0; i < rows; i++) { s->satdForVbv[i] = rowStat[i].satdForVbv; s->intraSatdForVbv[i] = rowStat[i].intraSatdForVbv; }
    if (p->rc.vbvBufferSize && p->rc.vbvMaxBitrate && l.lowresCostForRc)
        s->lowresCostForRc.assign(l.lowresCostForRc, l.lowresCostForRc + ncu);
    else
        s->lowresCostForRc.assign(ncu, 0);
    s->intraCostForRc.assign(l.intraCost, l.intraCost + ncu);
    {
        /* which estimate it read: the same derivation as the call itself */
        int d0 = 0, d1 = 0;
        if (slice.m_sliceType == P_SLICE) d0 = cur->m_poc - r0->m_poc;
        else if (slice.m_sliceType == B_SLICE) { d0 = r0 ? cur->m_poc - r0->m_poc : 0; d1 = r1->m_poc - cur->m_poc; }
        s->estRowSatds.assign(l.rowSatds[d0][d1], l.rowSatds[d0][d1] + l.maxBlocksInCol);
    }
    s->h.satdForVbv = s->satdForVbv.data(); s->h.intraSatdForVbv = s->intraSatdForVbv.data();
    s->h.lowresCostForRc = s->lowresCostForRc.data(); s->h.intraCostForRc = s->intraCostForRc.data();
    s->h.estRowSatds = s->estRowSatds.data();
    return 0;
}

int ref_la_num_out(void* hv) { return (int)((Handle*)hv)->out.size(); }
double ref_la_snapshot_seconds(void* hv) { return ((Handle*)hv)->secondsInSnapshot; }

/* free the arrays of snapshot `idx` (its scalars stay readable): a long full-size run is compared frame by frame */
void ref_la_drop(void* hv, int idx)
{
    Handle* h = (Handle*)hv;
    if (idx < 0 || idx >= (int)h->out.size()) return;
    FrameSnap* s = h->out[idx];
    FrameSnap* t = new FrameSnap;
    memset(&t->h, 0, sizeof(t->h));
    t->h.poc = s->h.poc; t->h.sliceType = s->h.sliceType; t->h.bScenecut = s->h.bScenecut; t->h.bKeyframe = s->h.bKeyframe;
    t->frame = s->frame;
    delete s;
    h->out[idx] = t;
}
double ref_la_seconds(void* hv) { return ((Handle*)hv)->secondsInLookahead; }

int ref_la_get(void* hv, int idx, RefLaFrame* out)
{
    Handle* h = (Handle*)hv;
    if (idx < 0 || idx >= (int)h->out.size()) return -1;
    *out = h->out[idx]->h;
    return 0;
}

void ref_la_close(void* hv)
{
    Handle* h = (Handle*)hv;
    for (size_t i = 0; i < h->out.size(); i++) delete h->out[i];
    /* Lookahead::destroy (via encoder close) frees frames still queued; drained ones are ours */
    h->la->stopJobs();
    for (size_t i = 0; i < h->retired.size(); i++) { h->retired[i]->destroy(); delete h->retired[i]; }
    x265_encoder_close(h->enc);
    PARAM_NS::x265_param_free(h->userParam);
    delete h;
}

/* ---- primitive-level taps: the reference's own C primitives on caller buffers ---- */

/* the lookahead's mvcost table row (encoder/bitcost.cpp:30-54) for X265_LOOKAHEAD_QP,
 * entries [-n..n] copied to out[0..2n] */
int ref_mvcost_table(uint16_t* out, int n)
{
    BitCost bc;
    bc.setQP(X265_LOOKAHEAD_QP);
    for (int i = -n; i <= n; i++) out[i + n] = bc.m_cost[i];
    return X265_LOOKAHEAD_QP;
}

int ref_lookahead_lambda(void) { return (int)x265_lambda_tab[X265_LOOKAHEAD_QP]; }

int ref_satd8x8(const void* a, int sa, const void* b, int sb)
{ return primitives.pu[LUMA_8x8].satd((const pixel*)a, sa, (const pixel*)b, sb); }
int ref_sad8x8(const void* a, int sa, const void* b, int sb)
{ return primitives.pu[LUMA_8x8].sad((const pixel*)a, sa, (const pixel*)b, sb); }
int ref_exp2fix8(double x) { return x265_exp2fix8(x); }

void ref_setup_primitives(void)
{
    x265_param* p = PARAM_NS::x265_param_alloc();
    PARAM_NS::x265_param_default(p);
    p->logLevel = X265_LOG_NONE;
    x265_setup_primitives(p);
    PARAM_NS::x265_param_free(p);
}

} // extern "C"
