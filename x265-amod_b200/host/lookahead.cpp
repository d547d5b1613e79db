/* lookahead.cpp -- host decision logic of the B200 lookahead (see lookahead.h).
 *
 * Behavioural reference: source/encoder/slicetype.cpp of DJATOM/x265-aMod 3.6+1.  Each routine
 * names the reference lines whose observable behaviour it reproduces.  The code is organised
 * around a publish/consume cache over GPU batches rather than around the reference's thread
 * pool, so only the decisions (and their order of first touch) are shared with it.
 *
 * Not supported (create() fails with an error string rather than silently diverging):
 * --hme, --hist-scenecut, aq-mode 4/5, hevc-aq, aq-motion,
 * zones, temporal sub-layers, analysis load, chunked encodes; --fades only with 16x16 AQ blocks and picture sizes whose
 * remainder modulo 16 is 0 or >= 8 (x265cu_create refuses the rest).
 */
#include "lookahead.h"
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <chrono>

namespace x265cu {

static inline double nowSec()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static inline double clipDuration(double f) { return f < 0.01 ? 0.01 : (f > 1.0 ? 1.0 : f); }  /* ratecontrol.h:43-48 */

/* x265_lambda_tab[X265_LOOKAHEAD_QP] (constants.cpp:34-90, common.h:209-213): exactly 1, 16, 256 */
int lookaheadLambda(int depth) { return 1 << (2 * (depth - 8)); }

/* The BitCost row the lookahead's MotionEstimate uses (bitcost.cpp:30-54, 98-113).  Built on the host like the reference and
 * uploaded; the device never recomputes it.  The reference writes `log((float)(i + 1)) * log2_2 + 1.718f` with <math.h>'s
 * log: the DOUBLE log of the float argument, times the float constant, plus the float constant, all in double, rounded to
 * float ONCE when it is stored in s_bitsizes.  (A float logf pipeline gives the same table for 8 and 10 bits but 15 entries
 * of +-60000 off by one at 12 bits, where lambda is 256 -- found by tools/fuzz_host_vs_reference.py.) */
void buildMvCostTable(std::vector<uint16_t>& table, int half, int depth)
{
    table.assign(2 * (size_t)half + 1, 0);
    const double lambda = (double)lookaheadLambda(depth);
    const float log2_2 = (float)(2.0 / log(2.0));
    for (int i = 0; i <= half; i++)
    {
        float bits = i ? (float)(log((double)(float)(i + 1)) * (double)log2_2 + (double)1.718f) : 0.718f;
        double c = bits * lambda + 0.5f;
        if (c > 32767.0) c = 32767.0;
        table[half + i] = table[half - i] = (uint16_t)c;
    }
}

void lookaheadParamDefault(LookaheadParam* p)
{
    memset(p, 0, sizeof(*p));
    p->internalBitDepth = 8; p->maxCUSize = 64;
    p->fpsNum = 25; p->fpsDenom = 1;
    p->bframes = 4; p->lookaheadDepth = 20; p->bFrameAdaptive = B_ADAPT_TRELLIS; p->bBPyramid = 1;
    p->scenecutThreshold = 40; p->scenecutBias = 0.05;
    p->keyframeMax = 250; p->keyframeMin = 0; p->bOpenGOP = 1;
    p->bEnableWeightedPred = 1; p->maxNumReferences = 3;
    p->rc.aqMode = 2; p->rc.aqStrength = 1.0; p->rc.cuTree = 1; p->rc.qCompress = 0.6; p->rc.qgSize = 32;
    p->rc.rateControlMode = 2 /* X265_RC_CRF */;
    p->extraSlots = 8; p->speculate = 1; p->asyncDepth = 0; p->pendingMax = 8;
    p->hmeSearchMethod[0] = 1; p->hmeSearchMethod[1] = 2; p->hmeRange[0] = 16; p->hmeRange[1] = 32;      /* param.cpp:226-230 */
}

Lookahead::Lookahead(const LookaheadParam& param)
    : m_param(param), m_filled(false), m_inputCount(0), m_lastNonB(NULL), m_lastNonBFrame(NULL),
      m_isSceneTransition(false), m_extendGopBoundary(false), m_rowsPerSlice(0), m_dualSlicing(false), m_inBatch(false),
      m_costVariants(2), m_ctx(NULL), m_pocNext(0), m_shardRank(0), m_failed(false)
{
    m_error[0] = 0;
    memset(m_timers, 0, sizeof(m_timers));
    /* slicetype.cpp:996-1033 */
    m_8x8Height = ((m_param.sourceHeight / 2) + 7) >> 3;
    m_8x8Width = ((m_param.sourceWidth / 2) + 7) >> 3;
    m_cuCount = m_8x8Width * m_8x8Height;
    m_8x8Blocks = m_8x8Width > 2 && m_8x8Height > 2 ? (m_cuCount + 4 - 2 * (m_8x8Width + m_8x8Height)) : m_cuCount;
    m_cuTreeStrength = 5.0 * (1.0 - m_param.rc.qCompress);
    if (!m_param.keyframeMin)   /* Encoder::configure, encoder.cpp:3658-3663 */
    {
        double fps = (double)m_param.fpsNum / m_param.fpsDenom;
        m_param.keyframeMin = std::min((int)fps, m_param.keyframeMax / 10);
    }
    m_param.keyframeMin = std::max(1, m_param.keyframeMin);
    if (m_param.gopLookahead && m_param.gopLookahead > m_param.lookaheadDepth - m_param.bframes - 2)   /* :1060-1064 */
        m_param.gopLookahead = std::max(0, m_param.lookaheadDepth - m_param.bframes - 2);
    m_lastKeyframe = -m_param.keyframeMax;
    m_shardDecouple = getenv("X265LA_SHARD_DECOUPLE") != NULL && atoi(getenv("X265LA_SHARD_DECOUPLE")) != 0;   /* opt-in: green on the gloo ranks, not yet timed on 2 GPUs */
    m_isFadeIn = false; m_fadeCount = 0; m_fadeStart = -1;      /* slicetype.cpp:1002-1004 */
    memset(m_accHistDiffRunningAvg, 0, sizeof(m_accHistDiffRunningAvg));        /* :1065-1095 */
    memset(m_accHistDiffRunningAvgCb, 0, sizeof(m_accHistDiffRunningAvgCb));
    memset(m_accHistDiffRunningAvgCr, 0, sizeof(m_accHistDiffRunningAvgCr));
    m_resetRunningAvg = true;
    m_segmentCountThreshold = (uint32_t)(((float)((4 * 4) * 50) / 100) + 0.5);
    for (int i = 0; i < BFRAME_MAX + 4; i++) m_frameVariance[i] = -1;
    /* asyncDepth extra frames of input delay: the decision only ever analyses the first rc-lookahead frames of the
     * queue (slicetype.cpp:1821-1827, 2609-2616), so the results are the same, but the GPU always holds that many
     * frames of searches in flight beyond the window being decided */
    m_fullQueueSize = std::max(1, m_param.lookaheadDepth) + std::max(0, m_param.asyncDepth);
    if (m_param.batchMin <= 0) m_param.batchMin = std::max(1, m_param.asyncDepth / 2);
    m_bAdaptiveQuant = m_param.rc.aqMode || m_param.bEnableWeightedPred || m_param.bEnableWeightedBiPred;
    m_bBatchMotionSearch = m_param.poolWorkers > 0 && m_param.bFrameAdaptive == B_ADAPT_TRELLIS;
    m_bBatchFrameCosts = m_bBatchMotionSearch;
    memset(&m_geom, 0, sizeof(m_geom));
}

Lookahead::~Lookahead() { destroy(); }

void Lookahead::fail(const char* what)
{
    if (!m_failed)
    {
        snprintf(m_error, sizeof(m_error), "%s%s%s", what, m_ctx ? ": " : "", m_ctx ? x265cu_last_error(m_ctx) : "");
        m_failed = true;
    }
}

bool Lookahead::check(int status, const char* what)
{
    if (status == X265CU_OK)
        return true;
    char buf[200];
    snprintf(buf, sizeof(buf), "%s failed (%s)", what, x265cu_strerror(status));
    fail(buf);
    return false;
}

bool Lookahead::create()
{
    const LookaheadParam& p = m_param;
    /* cooperative slices (slicetype.cpp:1035-1059).  They only exist with a thread pool and >= 720 lines.  A search
     * that is first needed outside one of the pool's batches runs sliced (:4004); without the batches (b-adapt 0 / 1)
     * that is every search, with them (b-adapt 2) both variants of a search can be needed and get their own stores. */
    int rowsPerSlice = 0;
    {
        int slices = p.lookaheadSlices;
        if (slices && p.poolWorkers <= 0) slices = 0;
        if (slices && p.sourceHeight < 720) slices = 0;
        if (slices > 1)
        {
            int rows = p.numRowsPerSlice > 0 ? p.numRowsPerSlice : std::min(std::max(m_8x8Height / slices, 10), m_8x8Height);
            if (m_8x8Height / rows > 1) rowsPerSlice = rows;
        }
    }
    m_rowsPerSlice = rowsPerSlice;
    m_dualSlicing = rowsPerSlice > 0 && m_bBatchMotionSearch;
    m_costVariants = m_dualSlicing ? 8 : 2;
    if (p.rc.qgSize != 8 && p.rc.qgSize != 16 && p.rc.qgSize != 32 && p.rc.qgSize != 64) { fail("qg-size must be 8, 16, 32 or 64"); return false; }
    if (p.rc.aqMode < 0 || p.rc.aqMode > 5) { fail("aq-mode must be 0..5"); return false; }
    if (p.rc.aqMode > 3 && p.bEnableFades)
    {
        /* the second acEnergyCu pass of --fades adds the luma sums to the weightp statistics once more, but not the edge image's
         * (slicetype.cpp:54-55, 568-573, 703): the statistics kernels keep one sum per plane */
        fail("--fades with aq-mode 4/5 (edge) is not supported by the GPU lookahead");
        return false;
    }
    if (p.bHistBasedSceneCut && p.internalBitDepth != 8) { fail("--hist-scenecut is 8-bit only (the reference indexes 256 bins with the sample value)"); return false; }
    if (p.bEnableTemporalSubLayers > 5) { fail("temporal-layers must be 0..5"); return false; }
    if (p.bEnableTemporalSubLayers > 2)
    {
        /* Encoder::configure (encoder.cpp:3914-3943) has fixed the mini-GOP to 3 / 7 / 15 B frames and turned b-adapt off */
        static const int tlBframes[6] = { 0, 0, 0, 3, 7, 15 };
        if (p.bframes != tlBframes[p.bEnableTemporalSubLayers] || p.bFrameAdaptive) { fail("temporal-layers 3..5 need bframes 3 / 7 / 15 and b-adapt 0 (Encoder::configure sets them)"); return false; }
        if (!p.lookaheadDepth) { fail("temporal-layers 3..5 with rc-lookahead 0 is not supported"); return false; }
    }
    if (p.bEnableHME && rowsPerSlice > 0)
    {
        /* the reference cuts the two levels into slices at different rows (slicetype.cpp:3942-3968): a slice's lowres search
         * reads level-0 vectors another worker may not have written yet (uninitialised memory; it crashes the reference here) */
        fail("--hme with active lookahead slices is undefined in the reference (its two levels race); use --lookahead-slices 0");
        return false;
    }
    if (p.bEnableHME)
        for (int i = 0; i < 2; i++)
        {
            if (p.hmeSearchMethod[i] < 0 || p.hmeSearchMethod[i] > 5 || p.hmeSearchMethod[i] == 4) { fail("--hme-search: sea is not supported for levels 0 and 1 by the GPU lookahead (dia, hex, umh, star, full are)"); return false; }
            if (p.hmeRange[i] < 4 || p.hmeRange[i] > 256) { fail("--hme-range of levels 0 and 1 must be 4..256"); return false; }
        }
    if (p.bframes > BFRAME_MAX || p.bframes < 0) { fail("bframes out of range"); return false; }
    if (p.lookaheadDepth && p.lookaheadDepth <= p.bframes) { fail("rc-lookahead must exceed bframes"); return false; }
    if (p.lookaheadDepth > LOOKAHEAD_MAX) { fail("rc-lookahead too large"); return false; }

    const int lowW = 8 * m_8x8Width;
    int half = std::min(2 * 32768, 4 * (lowW + 8 * m_8x8Height) * 2 + 4096);
    if (p.bEnableHME)   /* the star search's raster pass charges one candidate in four for the vector shifted by 3 bits (motion.cpp:1219) */
        half = std::min(2 * 32768, std::max(half, 12 * (std::max(lowW, 8 * m_8x8Height) + 16) + 1024));
    buildMvCostTable(m_mvcost, half, p.internalBitDepth);

    x265cu_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.width = p.sourceWidth; cfg.height = p.sourceHeight; cfg.depth = p.internalBitDepth;
    cfg.max_cu_size = p.maxCUSize; cfg.bframes = p.bframes;
    cfg.max_slots = std::max(1, p.lookaheadDepth) + 3 * (p.bframes + 2) + 4 + p.extraSlots + std::max(0, p.asyncDepth);
    cfg.qg_size = p.rc.qgSize; cfg.aq_mode = p.rc.aqMode; cfg.aq_strength = p.rc.aqStrength;
    cfg.need_aq = m_bAdaptiveQuant; cfg.need_wp_stats = p.bEnableWeightedPred || p.bEnableWeightedBiPred;
    cfg.fade_stats = p.bEnableFades; cfg.hist_stats = p.bHistBasedSceneCut;
    cfg.hme = p.bEnableHME;
    for (int i = 0; i < 2; i++) { cfg.hme_search[i] = p.hmeSearchMethod[i]; cfg.hme_range[i] = p.hmeRange[i]; }
    cfg.lambda = lookaheadLambda(p.internalBitDepth);
    cfg.mvcost = &m_mvcost[0]; cfg.mvcost_half = half;
    cfg.device = p.device;
    cfg.rows_per_slice = rowsPerSlice;
    cfg.mv_store_kinds = m_dualSlicing ? 6 : 3; cfg.cost_variants = m_costVariants;
    if (!check(x265cu_create(&cfg, &m_ctx), "x265cu_create"))
        return false;
    if (!check(x265cu_get_geometry(m_ctx, &m_geom), "x265cu_get_geometry"))
        return false;
    if (m_geom.bw != m_8x8Width || m_geom.bh != m_8x8Height) { fail("engine geometry mismatch"); return false; }
    m_pool.resize(cfg.max_slots);
    for (size_t i = 0; i < m_pool.size(); i++)
    {
        Frame* f = new Frame;
        memset(f, 0, sizeof(*f));
        f->m_lowres.slot = (int)i;
        f->m_owner = this;
        m_pool[i] = f;
    }
    return true;
}

void Lookahead::destroy()
{
    for (size_t i = 0; i < m_pool.size(); i++) delete m_pool[i];
    m_pool.clear(); m_inputQueue.clear(); m_outputQueue.clear(); m_resident.clear(); m_pendingSpec.clear(); m_unverified.clear();
    if (m_ctx)
    {
        x265cu_sync(m_ctx);
        for (std::map<const void*, bool>::iterator it = m_pinned.begin(); it != m_pinned.end(); ++it)
            if (it->second) x265cu_unpin_host(m_ctx, const_cast<void*>(it->first));
        m_pinned.clear();
        x265cu_destroy(m_ctx); m_ctx = NULL;
    }
}

/* Lowres::init minus the pixel work (common/lowres.cpp:337-365) */
void Lookahead::initLowres(Frame* f, int poc)
{
    Lowres& l = f->m_lowres;
    int slot = l.slot;
    memset(&l, 0, sizeof(l));
    l.slot = slot;
    l.frameNum = poc;
    l.satdCost = -1;
    l.rcD0 = l.rcD1 = -1;
    l.rcPlanD0 = l.rcPlanD1 = -1;
    for (int i = 0; i < BFRAME_MAX + 2; i++)
        for (int j = 0; j < BFRAME_MAX + 2; j++)
        {
            /* costEstAq is only reset to -1 when the AQ arrays exist (lowres.cpp:348-349) */
            l.costEst[i][j] = -1; l.costEstAq[i][j] = m_bAdaptiveQuant ? -1 : 0; l.costStore[i][j] = -1;
        }
    for (int i = 0; i < BFRAME_MAX + 2; i++)
        l.mvStore[0][i] = l.mvStore[1][i] = -1;
}

Frame* Lookahead::frameOfPoc(int poc)
{
    for (size_t i = m_resident.size(); i-- > 0;)
        if (m_resident[i]->m_poc == poc)
            return m_resident[i];
    return NULL;
}

/* Release slots nobody can reference any more: the caller is done with the frame, it is not
 * m_lastNonB, and it is more than bframes+1 behind the newest arrival (no future search or
 * estimate can name it, slicetype.cpp:2674-2689 / 3221). */
void Lookahead::recycle()
{
    size_t w = 0;
    for (size_t i = 0; i < m_resident.size(); i++)
    {
        Frame* f = m_resident[i];
        bool dead = f->m_released && f != m_lastNonBFrame && f->m_poc + m_param.bframes + 1 < m_pocNext - 1;
        if (dead) f->m_inUse = false;
        else m_resident[w++] = f;
    }
    m_resident.resize(w);
}

void Lookahead::releaseFrame(Frame* f) { if (f) { f->m_released = true; recycle(); } }

/* slicetype.cpp:1200-1243 */
Frame* Lookahead::addPicture(const void* y, const void* u, const void* v, int strideY, int strideC,
                             int64_t pts, int sliceType, int sliceTypeReq)
{
    if (m_failed) return NULL;
    if (!m_filled)   /* checkLookaheadQueue */
    {
        if (!m_param.bframes & !m_param.lookaheadDepth) m_filled = true;
        else if (m_inputCount >= m_param.lookaheadDepth + 2 + m_param.bframes + std::max(0, m_param.asyncDepth)) m_filled = true;
    }
    Frame* f = NULL;
    for (size_t i = 0; i < m_pool.size(); i++)
        if (!m_pool[i]->m_inUse) { f = m_pool[i]; break; }
    if (!f) { recycle(); for (size_t i = 0; i < m_pool.size(); i++) if (!m_pool[i]->m_inUse) { f = m_pool[i]; break; } }
    if (!f)
    {
        /* recoverable (the caller can release frames and retry), so the context is not marked failed */
        snprintf(m_error, sizeof(m_error), "no free frame slot: the caller holds too many unreleased frames (raise LookaheadParam::extraSlots)");
        return NULL;
    }
    f->m_inUse = true; f->m_released = false; f->m_speculated = false; f->m_lowresInit = false;
    f->m_poc = m_pocNext++; f->m_pts = pts; f->m_reorderedPts = 0;
    f->m_gopOffset = 0; f->m_gopId = 0; f->m_tempLayer = 0; f->m_gopIdSet = false;
    f->m_planes[0] = y; f->m_planes[1] = u; f->m_planes[2] = v; f->m_strideY = strideY; f->m_strideC = strideC;
    initLowres(f, f->m_poc);
    /* Encoder::encode (encoder.cpp:1713-1714, 1863): the type an application forces through x265_picture::sliceType goes
     * to Lowres::sliceTypeReq -- the frame is analysed as AUTO and the type re-imposed at slicetype.cpp:1938 --, while
     * addPicture's own argument is the first-pass type of a 2-pass encode and lands in Lowres::sliceType */
    f->m_lowres.sliceType = sliceType;
    f->m_lowres.sliceTypeReq = sliceTypeReq;
    if (m_param.shardCount > 1 &&
        !check(x265cu_slot_owner(m_ctx, f->m_lowres.slot, f->m_poc % m_param.shardCount), "x265cu_slot_owner"))
        return NULL;
    if (m_param.pinHost)
    {
        /* page-lock the caller's picture buffers the first time they are seen (an encoder recycles a fixed set of them,
         * PicYuv via the DPB free list, encoder.cpp:1632): pinned uploads run at full PCIe rate and truly asynchronously */
        const size_t bpp = m_param.internalBitDepth > 8 ? 2 : 1;
        const int cw = (m_param.sourceWidth + 1) / 2, ch = (m_param.sourceHeight + 1) / 2;
        const void* ptr[3] = { y, u, v };
        const size_t len[3] = { ((size_t)strideY * (m_param.sourceHeight - 1) + m_param.sourceWidth) * bpp,
                                ((size_t)strideC * (ch - 1) + cw) * bpp, ((size_t)strideC * (ch - 1) + cw) * bpp };
        for (int i = 0; i < 3; i++)
            if (ptr[i] && !m_pinned.count(ptr[i]))
            {
                /* best effort: a refusal (locked-memory limit) only costs speed */
                m_pinned[ptr[i]] = x265cu_pin_host(m_ctx, const_cast<void*>(ptr[i]), len[i]) == X265CU_OK;
            }
    }
    if (m_param.bHistBasedSceneCut && (!u || !v))
    {
        /* its picture statistics read the chroma planes (slicetype.cpp:1560-1640); the reference dereferences NULL on a 4:0:0 picture */
        fail("--hist-scenecut needs chroma planes (4:0:0 picture)");
        return NULL;
    }
    /* the device starts the frame's pre-lookahead now; results are collected in slicetypeDecide */
    if (!check(x265cu_frame_upload(m_ctx, f->m_lowres.slot, y, u, v, strideY, strideC), "x265cu_frame_upload"))
        return NULL;
    m_resident.push_back(f);
    m_inputQueue.push_back(f);
    m_inputCount++;
    if (m_param.speculate)
    {
        m_pendingSpec.push_back(f);
        if (m_param.speculate == 1 && m_param.shardCount <= 1 && (int)m_pendingSpec.size() >= std::min(m_param.batchMin, 8))
        {
            /* per-decision mode, window still filling (or the host far ahead of the GPU): hand the frames whose
             * pre-lookahead has finished to the GPU in launches of batchMin frames instead of one launch at the first decision */
            const double t0 = nowSec();
            drainPending((size_t)m_param.pendingMax, -1);
            m_timers[7] += nowSec() - t0;
            if (m_failed) return NULL;
        }
        if (m_param.speculate >= 2)
        {
            /* streaming: the frame's searches and costs go to the GPU now, long before a decision asks for them.
             * With weightp the analysis needs the frame's pixel sums on the host, so it runs one frame late rather
             * than wait for this frame's upload */
            const double t0 = nowSec();
            drainPending((size_t)m_param.pendingMax, -1);
            m_timers[7] += nowSec() - t0;
            if (m_failed) return NULL;
        }
    }
    return f;
}

void Lookahead::flush() { m_fullQueueSize = 1; m_filled = true; }   /* slicetype.cpp:1246-1251 */

/* slicetype.cpp:1259-1322 without the thread hand-off: the decision runs on the caller */
Frame* Lookahead::getDecidedPicture()
{
    if (!m_filled || m_failed)
        return NULL;
    /* The reference decides as soon as the input queue is full (a pool worker runs slicetypeDecide the moment addPicture
     * fills it, slicetype.cpp:1231-1243), so its output queue usually holds the mini-GOP decided one step earlier.  Same
     * here: decide while at most one mini-GOP is left in the output queue.  A caller that takes one decided frame per
     * picture it adds (Encoder::encode) then reads a frame's rate-control cost and mirrors one mini-GOP after its decision,
     * when the cuTree pass that produces them has long finished on the GPU, instead of waiting for it. */
    if ((int)m_outputQueue.size() <= m_param.bframes + 1 && (int)m_inputQueue.size() >= m_fullQueueSize && !m_inputQueue.empty())
        slicetypeDecide();
    if (m_outputQueue.empty() || m_failed)
        return NULL;
    Frame* out = m_outputQueue.front();
    m_outputQueue.pop_front();
    m_inputCount--;
    return out;
}

int Lookahead::findSliceType(int poc)   /* slicetype.cpp:3248-3266 */
{
    if (m_filled)
        for (size_t i = 0; i < m_outputQueue.size(); i++)
            if (m_outputQueue[i]->m_poc == poc)
                return m_outputQueue[i]->m_lowres.sliceType;
    return TYPE_AUTO;
}

/* -------------------------------------------------------------------------------------------
 * device orchestration
 * ------------------------------------------------------------------------------------------- */

/* PreLookaheadGroup::processTasks (slicetype.cpp:1726-1752): the pixel work was enqueued by
 * addPicture; collect the scalars the decisions need. */
void Lookahead::preLookahead(const std::vector<Frame*>& fr)
{
    if (fr.empty()) return;
    std::vector<int32_t> slots(fr.size());
    std::vector<x265cu_frame_stats> st(fr.size());
    for (size_t i = 0; i < fr.size(); i++) slots[i] = fr[i]->m_lowres.slot;
    if (!check(x265cu_frame_stats_get(m_ctx, &slots[0], (int)slots.size(), &st[0]), "x265cu_frame_stats_get"))
        return;
    for (size_t i = 0; i < fr.size(); i++)
    {
        Lowres& l = fr[i]->m_lowres;
        l.costEst[0][0] = st[i].cost_est;
        l.costEstAq[0][0] = st[i].cost_est_aq;
        l.rowSatdsValid[0][0] = true;
        l.costStore[0][0] = 0;
        for (int k = 0; k < 3; k++) { l.wp_ssd[k] = st[i].wp_ssd[k]; l.wp_sum[k] = st[i].wp_sum[k]; }
        l.frameVariance = st[i].frame_variance;
        if (m_param.bHistBasedSceneCut)
        {
            fr[i]->m_hist.resize(1);
            if (!check(x265cu_frame_hist_get(m_ctx, l.slot, &fr[i]->m_hist[0]), "x265cu_frame_hist_get")) return;
            l.hist = &fr[i]->m_hist[0];
            l.bHistScenecutAnalyzed = false;        /* collectPictureStatistics, :1723 */
        }
        l.statsFetched = true;
        fr[i]->m_lowresInit = true;
    }
}

/* LookaheadTLD::weightsAnalyse (slicetype.cpp:879-980) for many (fenc, ref) pairs at once.  The
 * scalar logic is the reference's; the two whole-frame SATD scores per pair are GPU batches. */
void Lookahead::weightsAnalyseBatch(const std::vector<std::pair<Lowres*, Lowres*> >& pairs)
{
    struct Cand { Lowres* fenc; Lowres* ref; int d; int mindenom, minscale, curScale, curOffset; unsigned orig; float fencMean, refMean; };
    std::vector<Cand> cands;
    const int depthShift = m_param.internalBitDepth - 8;
    const int lowW = m_geom.low_width, lowH = m_geom.low_height;
    for (size_t i = 0; i < pairs.size(); i++)
    {
        Lowres* fenc = pairs[i].first; Lowres* ref = pairs[i].second;
        int d = fenc->frameNum - ref->frameNum;
        if (fenc->weightState[d]) continue;
        fenc->weightState[d] = 1;
        const float epsilon = 1.f / 128.f;
        float guessScale, fencMean, refMean;
        if (fenc->wp_ssd[0] && ref->wp_ssd[0])
            guessScale = sqrtf((float)fenc->wp_ssd[0] / ref->wp_ssd[0]);
        else
            guessScale = 1.0f;
        fencMean = (float)fenc->wp_sum[0] / (lowH * lowW) / (1 << depthShift);
        refMean = (float)ref->wp_sum[0] / (lowH * lowW) / (1 << depthShift);
        if (fabsf(refMean - fencMean) < 0.5f && fabsf(1.f - guessScale) < epsilon)
            continue;
        Cand c; c.fenc = fenc; c.ref = ref; c.d = d; c.fencMean = fencMean; c.refMean = refMean;
        int w = (int)(guessScale * 128 + 0.5f), den = 7;     /* setFromWeightAndOffset(w,0,7,true), slice.h:304-316 */
        while (den > 0 && w > 127) { den--; w >>= 1; }
        c.mindenom = den; c.minscale = std::min(w, 127);
        c.curScale = c.curOffset = 0; c.orig = 0;
        cands.push_back(c);
    }
    if (cands.empty()) return;
    std::vector<x265cu_wcost_job> jobs(cands.size());
    std::vector<uint32_t> costs(cands.size());
    for (size_t i = 0; i < cands.size(); i++)
    {
        memset(&jobs[i], 0, sizeof(jobs[i]));
        jobs[i].fenc_slot = cands[i].fenc->slot; jobs[i].ref_slot = cands[i].ref->slot;
    }
    if (!check(x265cu_weight_cost_batch(m_ctx, &jobs[0], (int)jobs.size(), &costs[0]), "x265cu_weight_cost_batch"))
        return;
    std::vector<Cand> second;
    for (size_t i = 0; i < cands.size(); i++)
    {
        Cand c = cands[i];
        c.orig = costs[i];
        if (!c.orig) continue;
        c.curScale = c.minscale;
        c.curOffset = (int)(c.fencMean - c.refMean * c.curScale / (1 << c.mindenom) + 0.5f);
        if (c.curOffset < -128 || c.curOffset > 127)
        {
            c.curOffset = std::max(-128, std::min(127, c.curOffset));
            c.curScale = (int)((1 << c.mindenom) * (c.fencMean - c.curOffset) / c.refMean + 0.5f);
            c.curScale = std::max(0, std::min(127, c.curScale));
        }
        second.push_back(c);
    }
    if (second.empty()) return;
    jobs.resize(second.size()); costs.resize(second.size());
    for (size_t i = 0; i < second.size(); i++)
    {
        jobs[i].fenc_slot = second[i].fenc->slot; jobs[i].ref_slot = second[i].ref->slot;
        jobs[i].weighted = 1; jobs[i].w_scale = second[i].curScale; jobs[i].w_denom = second[i].mindenom;
        jobs[i].w_offset = second[i].curOffset;
    }
    if (!check(x265cu_weight_cost_batch(m_ctx, &jobs[0], (int)jobs.size(), &costs[0]), "x265cu_weight_cost_batch"))
        return;
    for (size_t i = 0; i < second.size(); i++)
    {
        Cand& c = second[i];
        unsigned minscore = c.orig, s = costs[i];
        int minscale = c.minscale, mindenom = c.mindenom, minoff = 0, found = 0;
        if (s < minscore) { minscore = s; minscale = c.curScale; minoff = c.curOffset; found = 1; }
        if (mindenom > 0 && !(minscale & 1))
        {
            int idx = 0;
            while (minscale && !((minscale >> idx) & 1)) idx++;
            if (!minscale) idx = 32;
            int shift = std::min(idx, mindenom);
            mindenom -= shift; minscale >>= shift;
        }
        if (!found || (minscale == (1 << mindenom) && minoff == 0) || (float)minscore / c.orig > 0.998f)
            continue;
        c.fenc->weightedCostDelta[c.d] = (double)(minscore / c.orig);   /* integer division, slicetype.cpp:964 */
        c.fenc->weightState[c.d] = 2;
        c.fenc->wScale[c.d] = minscale; c.fenc->wDenom[c.d] = mindenom; c.fenc->wOffset[c.d] = minoff;
    }
}

void Lookahead::addSearch(Lowres* fenc, Lowres* ref, int kind, int d, int condStore)
{
    if (fenc->haveSearch[kind][d]) return;
    fenc->haveSearch[kind][d] = 1;
    const int nb = m_geom.nb;
    x265cu_search_job j;
    memset(&j, 0, sizeof(j));
    j.fenc_slot = fenc->slot; j.ref_slot = ref->slot;
    j.bidir_ctx = kind % 3 != 0;
    j.store = kind * nb + d;
    j.cond_store = condStore;
    j.sliced = jobSliced(kind);
    if (kind % 3 < 2 && fenc->weightState[d] == 2)
    {
        j.weighted = 1; j.w_scale = fenc->wScale[d]; j.w_denom = fenc->wDenom[d]; j.w_offset = fenc->wOffset[d];
    }
    m_searchJobs.push_back(j);
}

void Lookahead::addCost(Lowres* b, Lowres* p0, Lowres* p1, int d0, int d1, int l0kind, int l1kind, int condStore)
{
    const int nb = m_geom.nb, variant = costVariant(l0kind, p1 ? l1kind : -1);
    if (b->haveCost[d0][d1][variant]) return;
    b->haveCost[d0][d1][variant] = 1;
    x265cu_cost_job j;
    j.b_slot = b->slot; j.p0_slot = p0->slot; j.p1_slot = p1 ? p1->slot : b->slot;
    j.l0_store = l0kind * nb + d0;
    j.l1_store = p1 ? l1kind * nb + d1 : -1;
    j.out = costStoreOf(d0, d1, variant);
    j.cond_store = condStore;
    m_costJobs.push_back(j);
}

void Lookahead::launchJobs()
{
    if (!m_searchJobs.empty())
        check(x265cu_search_batch(m_ctx, &m_searchJobs[0], (int)m_searchJobs.size()), "x265cu_search_batch");
    if (!m_costJobs.empty())
        check(x265cu_cost_batch(m_ctx, &m_costJobs[0], (int)m_costJobs.size()), "x265cu_cost_batch");
    m_searchJobs.clear(); m_costJobs.clear();
}

/* every frame cost reading the L0 searches of kind `l0kind` whose searches exist on the device and whose
 * frames are resident: P estimates (d0, 0) and B estimates (d0, d1) the reference can ask for (p1 - p0 <= bframes+1,
 * slicetype.cpp:3221-3305; every (d0, d1) when the frame-cost batches of :2696-2735 are emulated).
 * conditional: the jobs read a P-context L0 search that only exists if its B-context twin applied the skip rule */
void Lookahead::enqueueCosts(int l0kind, bool conditional)
{
    const int B = m_param.bframes, nb = m_geom.nb;
    const int l1kind = 2 + 3 * (l0kind / 3);        /* speculation pairs searches of the same slicedness */
    for (size_t i = 0; i < m_resident.size(); i++)
    {
        Frame* bf = m_resident[i];
        Lowres* b = &bf->m_lowres;
        for (int d0 = 1; d0 <= B + 1; d0++)
        {
            if (!b->haveSearch[l0kind][d0]) continue;
            Frame* p0f = frameOfPoc(bf->m_poc - d0);
            if (!p0f) continue;
            /* only a P-context search that was itself queued as the conditional twin of a B-context search
             * (haveSearch == 1) may be missing; one issued on demand (== 2) always exists */
            const int cond = (conditional && b->haveSearch[l0kind][d0] == 1) ? (l0kind + 1) * nb + d0 : -1;
            addCost(b, &p0f->m_lowres, NULL, d0, 0, l0kind, -1, cond);
            const int maxD1 = m_bBatchFrameCosts ? B : B + 1 - d0;
            for (int d1 = 1; d1 <= B + 1; d1++)
            {
                /* (d0, d0): the estimate a search batch makes while it searches both lists at distance d0 (:2686-2692) */
                if (d1 > maxD1 && !(m_bBatchMotionSearch && d1 == d0)) continue;
                if (!b->haveSearch[l1kind][d1]) continue;
                Frame* p1f = frameOfPoc(bf->m_poc + d1);
                if (!p1f) continue;
                addCost(b, &p0f->m_lowres, &p1f->m_lowres, d0, d1, l0kind, l1kind, cond);
            }
        }
    }
}

/* Eager speculation: for every frame in `fresh` (in arrival order), every motion search and frame cost the
 * reference could ask for (distances <= bframes+1, slicetype.cpp:2674-2689, 3221-3305) whose frames are
 * resident, as ONE asynchronous batch on the engine.  Nothing here waits for the GPU.
 * With B frames an L0 search exists in two variants (P / B context, see lookahead.h).  The B-context variant is
 * always computed; the P-context variant and the costs that read it are enqueued as CONDITIONAL jobs which the
 * device skips when the B-context search never applied the zero-MV skip rule (the two variants are then the same
 * search and the host aliases them once it has read the flag, resolveAlias). */
void Lookahead::speculateFrames(const std::vector<Frame*>& fresh, int onlyDist)
{
    if (fresh.empty() || m_failed) return;
    const int B = m_param.bframes, nb = m_geom.nb;
    double t0 = nowSec();
    if (!check(x265cu_batch_begin(m_ctx, NULL), "x265cu_batch_begin")) return;
    m_searchJobs.clear(); m_costJobs.clear();
    /* dual slicing: while the reference's pool batches run, (nearly) every search is first touched by one of them, i.e.
     * unsliced; once they are switched off (small pool, :2691) every search is first touched on demand, i.e. sliced.
     * The other variant is computed on demand if the control flow ever asks for it */
    const int s3 = (m_dualSlicing && !m_bBatchMotionSearch) ? 3 : 0;
    const int firstKind = (B > 0 ? 1 : 0) + s3;
    for (size_t i = 0; i < fresh.size(); i++)
    {
        Frame* fn = fresh[i];
        Lowres* n = &fn->m_lowres;
        fn->m_speculated = true;
        for (int d = 1; d <= B + 1; d++)
        {
            Frame* rf = frameOfPoc(fn->m_poc - d);
            if (!rf) break;
            if (onlyDist && d != onlyDist) continue;        /* verifyWeights redoes one L0 distance of a frame */
            Lowres* r = &rf->m_lowres;
            addSearch(n, r, firstKind, d);                  /* L0(n,d) */
            /* the reference's search batches pair every list-0 search with the list-1 search at the SAME distance
             * (p1 = b + i, slicetype.cpp:2686-2690), i.e. up to bframes + 1; on demand only distances <= bframes occur */
            if (B > 0 && (d <= B || m_bBatchMotionSearch))
                addSearch(r, n, 2 + s3, d);                 /* L1(n-d,d), reference = n */
        }
    }
    enqueueCosts(firstKind, false);
    launchJobs();
    if (B > 0 && !m_failed)
    {
        for (size_t i = 0; i < fresh.size(); i++)
        {
            Frame* fn = fresh[i];
            for (int d = 1; d <= B + 1; d++)
            {
                Frame* rf = frameOfPoc(fn->m_poc - d);
                if (!rf) break;
                if (onlyDist && d != onlyDist) continue;
                addSearch(&fn->m_lowres, &rf->m_lowres, s3, d, (1 + s3) * nb + d);
            }
        }
        enqueueCosts(s3, true);
        launchJobs();
    }
    check(x265cu_batch_end(m_ctx), "x265cu_batch_end");
    m_timers[2] += nowSec() - t0;
}

/* Speculate the frames that arrived but were not processed yet, oldest first.  Frames up to mustPoc are needed now;
 * newer ones are taken too -- the GPU works on them while the host decides from the results of earlier batches.
 * Nothing here waits for the GPU on behalf of a frame nobody needs yet.  In particular weightp: whether an L0 search runs
 * on a weighted reference is decided on the host from the pixel sums of the two frames (weightsAnalyse), and those
 * only exist once the frame's pre-lookahead has run -- which, behind a saturated search launch, is milliseconds
 * away.  A pair whose sums are not on the host yet is therefore enqueued ASSUMING no weight (weightState 3) and
 * verified later (verifyWeights); outside fades the assumption always holds.  Waiting for the sums instead (the first
 * version did) serialised the whole pipeline: batch k+1 could only be cut after the search launch of batch k had
 * drained.  Scheduling only: what is computed does not depend on when.
 * streaming mode: one batch per frame; otherwise one batch for all of them */
void Lookahead::drainPending(size_t keep, int mustPoc)
{
    std::vector<Frame*> group;
    const bool needStats = m_param.bEnableWeightedPred != 0;
    const bool sharded = m_param.shardCount > 1;
    (void)keep;
    /* per-decision mode: unless the window needs one of them now, wait until enough frames have gathered for a launch
     * that fills the GPU (a lowres search is a ~500-step wavefront: small launches spend most of their time ramping
     * up and draining) */
    if (m_param.speculate == 1 && !m_pendingSpec.empty() && m_pendingSpec.front()->m_poc > mustPoc)
    {
        /* ... unless the GPU is about to run dry: it should always hold the batch it is working on AND one queued behind
         * it (whose search launch ramps up while the first one drains), so with fewer than two in flight a smaller launch now
         * beats a full one later.  A sharded stream must batch the same frames on every rank, so it never asks. */
        const int minFrames = std::min(m_param.batchMin, 8);
        const bool hungry = !sharded && (int)m_pendingSpec.size() >= minFrames && (int)m_pendingSpec.size() < m_param.batchMin &&
                            x265cu_batches_in_flight(m_ctx) < 2;
        const int want = hungry ? minFrames : m_param.batchMin;
        if ((int)m_pendingSpec.size() < want)
            return;
    }
    /* weights assumed earlier whose pixel sums have arrived in the meantime: settle them first, so a search that does
     * need weights is redone before more work piles up behind the wrong one */
    const bool lockstep = sharded && !m_shardDecouple;      /* the first sharded pipeline: wait for every frame's sums before batching it */
    if (needStats && !lockstep) verifyWeights(mustPoc);
    while (!m_pendingSpec.empty() && !m_failed)
    {
        Frame* f = m_pendingSpec.front();
        /* (a sharded stream must batch the same frames on every rank, so it never looks at the clock: it leaves the
         * two newest frames, whose pre-lookahead is probably still running, for the next batch and waits for the sums) */
        if (f->m_poc > mustPoc && lockstep && needStats && f->m_poc > m_pocNext - 3)
            break;
        m_pendingSpec.pop_front();
        group.push_back(f);
    }
    if (group.empty()) return;
    if (needStats)
    {
        /* the pixel sums of the group's frames and of the references they pair with: waited for where a decision needs the
         * frame now (or the stream is sharded), otherwise taken only if they are already there */
        std::vector<Frame*> pre;
        for (size_t i = 0; i < group.size(); i++)
        {
            const bool must = lockstep || group[i]->m_poc <= mustPoc;
            for (int d = 0; d <= m_param.bframes + 1; d++)
            {
                Frame* x = d ? frameOfPoc(group[i]->m_poc - d) : group[i];
                if (!x) break;
                if (x->m_lowresInit || std::find(pre.begin(), pre.end(), x) != pre.end()) continue;
                /* (a sharded stream never asks the clock: ranks must make identical choices) */
                if (must || (!sharded && x265cu_frame_ready(m_ctx, x->m_lowres.slot) == 1))
                    pre.push_back(x);
            }
        }
        const double t0 = nowSec();
        preLookahead(pre);
        m_timers[0] += nowSec() - t0;
        /* the weightp analysis of the whole group in two round trips instead of two per frame */
        const double t1 = nowSec();
        std::vector<std::pair<Lowres*, Lowres*> > pairs;
        for (size_t i = 0; i < group.size(); i++)
            for (int d = 1; d <= m_param.bframes + 1; d++)
            {
                Frame* r = frameOfPoc(group[i]->m_poc - d);
                if (!r) break;
                if (group[i]->m_lowresInit && r->m_lowresInit)
                    pairs.push_back(std::make_pair(&group[i]->m_lowres, &r->m_lowres));
                else if (!group[i]->m_lowres.weightState[d])
                {
                    group[i]->m_lowres.weightState[d] = 3;
                    m_unverified.push_back(std::make_pair(group[i], d));
                }
            }
        weightsAnalyseBatch(pairs);
        m_timers[1] += nowSec() - t1;
    }
    if (m_param.speculate >= 2)
        for (size_t i = 0; i < group.size() && !m_failed; i++)
            speculateFrames(std::vector<Frame*>(1, group[i]));
    else
        speculateFrames(group);
}

/* Settle the weightp assumptions of drainPending: every (frame, L0 distance) enqueued without weights because the pixel
 * sums of the pair were still on their way.  A pair is analysed as soon as both frames' sums are on the host (taken
 * without waiting when their pre-lookahead has finished; waited for when the frame is <= mustPoc, i.e. a decision is about to
 * read its results).  weightsAnalyse's answer is "no weight" outside fades and nothing else happens; otherwise the L0
 * searches of that distance and every cost that read them are enqueued again, this time on the weighted reference
 * (the engine orders the new batch behind the old writers and readers of those stores).  No result of such a frame has
 * been read by the host before this point: results are only fetched for frames <= maxPoc of a decision, after this ran. */
void Lookahead::verifyWeights(int mustPoc)
{
    if (m_unverified.empty() || m_failed) return;
    std::vector<Frame*> need;
    for (size_t i = 0; i < m_unverified.size(); i++)
    {
        Frame* f = m_unverified[i].first;
        Frame* pair[2] = { f, frameOfPoc(f->m_poc - m_unverified[i].second) };
        for (int k = 0; k < 2; k++)
        {
            Frame* x = pair[k];
            if (!x || x->m_lowresInit || std::find(need.begin(), need.end(), x) != need.end()) continue;
            /* a sharded stream settles every assumption made by EARLIER batches at every call (their sums are at most one search
             * launch away), identically on all ranks; an unsharded one takes what has arrived */
            if (f->m_poc <= mustPoc || m_param.shardCount > 1 || x265cu_frame_ready(m_ctx, x->m_lowres.slot) == 1)
                need.push_back(x);
        }
    }
    if (!need.empty())
    {
        const double t0 = nowSec();
        preLookahead(need);
        m_timers[0] += nowSec() - t0;
        if (m_failed) return;
    }
    const double t1 = nowSec();
    std::vector<std::pair<Frame*, int> > later, checked;
    std::vector<std::pair<Lowres*, Lowres*> > pairs;
    for (size_t i = 0; i < m_unverified.size(); i++)
    {
        Frame* f = m_unverified[i].first;
        const int d = m_unverified[i].second;
        Frame* r = frameOfPoc(f->m_poc - d);
        if (!r)
        {
            /* The reference frame has left (decided, handed out and released) before this pair was settled: with an rc-lookahead not
             * much larger than bframes a frame is speculated against references up to bframes + 1 back that the decisions have already
             * passed.  No estimate can name that pair any more (frames[] only holds the last non-B and the queue), so the speculated
             * search is simply never read.  (Found by tools/fuzz_host_vs_reference.py; this used to fail the stream.) */
            continue;
        }
        if (f->m_lowresInit && r->m_lowresInit)
        {
            f->m_lowres.weightState[d] = 0;
            pairs.push_back(std::make_pair(&f->m_lowres, &r->m_lowres));
            checked.push_back(m_unverified[i]);
        }
        else
            later.push_back(m_unverified[i]);
    }
    m_unverified.swap(later);
    weightsAnalyseBatch(pairs);
    m_timers[1] += nowSec() - t1;
    if (m_failed) return;
    const int nb = m_geom.nb;
    for (size_t i = 0; i < checked.size() && !m_failed; i++)
    {
        Frame* f = checked[i].first;
        const int d = checked[i].second;
        Lowres& n = f->m_lowres;
        if (n.weightState[d] != 2) continue;
        /* the assumption was wrong: forget the L0 searches of this distance (both contexts, both slicednesses) and every
         * cost that read them, and enqueue them again */
        for (int kind = 0; kind < 6; kind++)
            if (kind % 3 < 2) n.haveSearch[kind][d] = 0;
        for (int s2 = 0; s2 < 2; s2++) { n.flagFetched[s2][d] = 0; n.l0Alias[s2][d] = 0; }
        for (int d1 = 0; d1 < nb; d1++)
            for (int v = 0; v < 8; v++)
            {
                if (n.resultFetched[d][d1][v]) { fail("verifyWeights: a result of an unverified search was already read"); return; }
                n.haveCost[d][d1][v] = 0;
            }
        if (getenv("X265LA_DEBUG_WEIGHTS")) fprintf(stderr, "verifyWeights: redo poc %d d %d\n", f->m_poc, d);
        speculateFrames(std::vector<Frame*>(1, f), d);
    }
}

/* the skip flags of the B-context L0 searches of `who` that the host has not looked at yet: decides, per (frame,
 * distance), whether the P-context variant is an alias (flag clear) or was really computed (flag set).  One
 * synchronising gather; waits only for the batches that ran those searches */
void Lookahead::resolveAlias(const std::vector<Lowres*>& who)
{
    const int B = m_param.bframes, nb = m_geom.nb;
    if (B <= 0 || m_failed) return;
    std::vector<int32_t> slots, stores, flags;
    struct Ref { Lowres* l; int s, d; };
    std::vector<Ref> ref;
    for (size_t i = 0; i < who.size(); i++)
        for (int s = 0; s <= (m_dualSlicing ? 1 : 0); s++)
            for (int d = 1; d <= B + 1; d++)
                if (who[i]->haveSearch[1 + 3 * s][d] && !who[i]->flagFetched[s][d])
                {
                    slots.push_back(who[i]->slot); stores.push_back((1 + 3 * s) * nb + d);
                    Ref r = { who[i], s, d };
                    ref.push_back(r);
                }
    if (slots.empty()) return;
    flags.resize(slots.size());
    if (!check(x265cu_search_flags_get(m_ctx, &slots[0], &stores[0], (int)slots.size(), &flags[0]), "x265cu_search_flags_get"))
        return;
    for (size_t i = 0; i < ref.size(); i++)
    {
        Lowres* l = ref[i].l; const int s = ref[i].s, d = ref[i].d;
        l->flagFetched[s][d] = 1;
        /* a P-context search issued unconditionally (demand path) is always the real thing */
        if (!l->l0Alias[s][d])
            l->l0Alias[s][d] = (flags[i] || l->haveSearch[3 * s][d] == 2) ? 2 : 1;
    }
}

/* one synchronising gather of every computed-but-unread cost scalar of `who` whose frames are all <= maxPoc (later
 * ones may still be in flight and no decision can ask for them yet) */
void Lookahead::fetchResults(const std::vector<Lowres*>& who, int maxPoc)
{
    const int nb = m_geom.nb;
    std::vector<int32_t> slots, outs;
    struct Ref { Lowres* l; int d0, d1, v; };
    std::vector<Ref> refs;
    for (size_t i = 0; i < who.size(); i++)
        for (int d0 = 0; d0 < nb; d0++)
            for (int d1 = 0; d1 < nb; d1++)
                for (int v = 0; v < m_costVariants; v++)
                    if (who[i]->haveCost[d0][d1][v] && !who[i]->resultFetched[d0][d1][v])
                    {
                        if (who[i]->frameNum + d1 > maxPoc) continue;
                        /* P-context L0 aliased to the B-context search: the conditional job did not run */
                        if ((v & 1) == 0 && who[i]->l0Alias[(v >> 1) & 1][d0] == 1) continue;
                        slots.push_back(who[i]->slot); outs.push_back(costStoreOf(d0, d1, v));
                        Ref r = { who[i], d0, d1, v };
                        refs.push_back(r);
                    }
    if (slots.empty()) return;
    std::vector<x265cu_cost_result> res(slots.size());
    if (!check(x265cu_cost_results_get(m_ctx, &slots[0], &outs[0], (int)slots.size(), &res[0]), "x265cu_cost_results_get"))
        return;
    for (size_t i = 0; i < refs.size(); i++)
    {
        refs[i].l->result[refs[i].d0][refs[i].d1][refs[i].v] = res[i];
        refs[i].l->resultFetched[refs[i].d0][refs[i].d1][refs[i].v] = 1;
    }
}

/* demand path: make sure the searches and the cost of one estimate exist on the device and its
 * scalars on the host (everything already speculated is a no-op) */
void Lookahead::ensureEstimate(Lowres* fenc, Lowres* ref0, Lowres* ref1, int d0, int d1, int l0kind, int l1kind)
{
    if (fenc->resultFetched[d0][d1][costVariant(l0kind, ref1 ? l1kind : -1)]) return;
    m_searchJobs.clear(); m_costJobs.clear();
    if (!fenc->haveSearch[l0kind][d0])
    {
        if (m_param.bEnableWeightedPred && !fenc->weightState[d0])
        {
            std::vector<std::pair<Lowres*, Lowres*> > one(1, std::make_pair(fenc, ref0));
            weightsAnalyseBatch(one);
        }
        addSearch(fenc, ref0, l0kind, d0);
        if (l0kind % 3 == 0) fenc->haveSearch[l0kind][d0] = 2;       /* unconditional */
    }
    if (ref1 && !fenc->haveSearch[l1kind][d1])
        addSearch(fenc, ref1, l1kind, d1);
    addCost(fenc, ref0, ref1, d0, d1, l0kind, l1kind);
    if (!check(x265cu_batch_begin(m_ctx, NULL), "x265cu_batch_begin")) return;
    launchJobs();
    check(x265cu_batch_end(m_ctx), "x265cu_batch_end");
    std::vector<Lowres*> who(1, fenc);
    fetchResults(who, 0x7fffffff);
}

/* -------------------------------------------------------------------------------------------
 * CostEstimateGroup::singleCost / estimateFrameCost (slicetype.cpp:3882-3886, 3976-4075)
 * ------------------------------------------------------------------------------------------- */

int64_t Lookahead::singleCost(Lowres** frames, int p0, int p1, int b, bool bIntraPenalty)
{
    return estimateFrameCost(frames, p0, p1, b, bIntraPenalty);
}

int64_t Lookahead::estimateFrameCost(Lowres** frames, int p0, int p1, int b, bool bIntraPenalty)
{
    Lowres* fenc = frames[b];
    const int d0 = b - p0, d1 = p1 - b, nb = m_geom.nb;
    int64_t score = 0;
    if (m_failed) return 0;

    if (fenc->costEst[d0][d1] >= 0 && fenc->rowSatdsValid[d0][d1])
        score = fenc->costEst[d0][d1];
    else
    {
        /* d0 == 0 with d1 > 0: a leading picture in front of a RADL IDR has no list-0 reference and the reference
         * estimates it against itself (slicetype.cpp:2385-2392 with p0 = b) */
        if (d0 < 0 || (d0 == 0 && d1 <= 0) || d0 >= nb || d1 < 0 || d1 >= nb) { fail("estimate outside the (bframes+2) window"); return 0; }
        const bool bDoSearch0 = fenc->mvStore[0][d0] < 0;
        const bool bDoSearch1 = p1 > b && fenc->mvStore[1][d1] < 0;
        /* first touch decides which variant of a search the reference would hold: its context (P / B estimate) and,
         * with cooperative slices next to pool batches, whether a batch or an on-demand estimate got there first */
        const int s3 = 3 * sliceNow();
        int l0kind = bDoSearch0 ? (p1 > b ? 1 : 0) + s3 : fenc->mvStore[0][d0] / nb;
        const int l1kind = p1 > b ? (bDoSearch1 ? 2 + s3 : fenc->mvStore[1][d1] / nb) : -1;
        if (l0kind % 3 == 0 && fenc->haveSearch[l0kind + 1][d0] && !fenc->flagFetched[l0kind / 3][d0])
            resolveAlias(std::vector<Lowres*>(1, fenc));
        l0kind = effKind(fenc, d0, l0kind);      /* identical variants share one store */
        ensureEstimate(fenc, frames[p0], p1 > b ? frames[p1] : NULL, d0, d1, l0kind, l1kind);
        if (m_failed) return 0;
        if (bDoSearch0) fenc->mvStore[0][d0] = l0kind * nb + d0;
        if (bDoSearch1) fenc->mvStore[1][d1] = l1kind * nb + d1;
        const int variant = costVariant(l0kind, l1kind);
        const x265cu_cost_result& r = fenc->result[d0][d1][variant];
        fenc->costEstAq[d0][d1] = r.cost_est_aq;
        if (p1 == b) fenc->intraMbs[d0] += r.intra_mbs;
        fenc->rowSatdsValid[d0][d1] = true;
        fenc->costStore[d0][d1] = costStoreOf(d0, d1, variant);
        score = r.cost_est;
        if (b != p1)
            score = score * 100 / (130 + m_param.bFrameBias);
        fenc->costEst[d0][d1] = score;
    }
    if (bIntraPenalty)
        score += score * fenc->intraMbs[b - p0] / (m_8x8Blocks * 8);
    return score;
}

/* -------------------------------------------------------------------------------------------
 * slicetypeDecide (slicetype.cpp:1802-2508; non-temporal-layer, non-analysis-load branches)
 * ------------------------------------------------------------------------------------------- */

void Lookahead::placeBref(Frame** list, int start, int end, int num, int* brefs)   /* :1755-1777 */
{
    int avg = (start + end) / 2;
    if (m_param.bEnableTemporalSubLayers < 2)
    {
        list[avg]->m_lowres.sliceType = TYPE_BREF;
        (*brefs)++;
        return;
    }
    if (num <= 2)
        return;
    list[avg]->m_lowres.sliceType = TYPE_BREF;
    (*brefs)++;
    placeBref(list, start, avg, avg - start, brefs);
    placeBref(list, avg + 1, end, end - avg, brefs);
}

/* :1780-1799 -- with two temporal layers the costs rate control will ask for are those of the B-ref hierarchy */
void Lookahead::compCostBref(Lowres** frames, int start, int end, int num)
{
    int avg = (start + end) / 2;
    if (num <= 2)
    {
        for (int i = start; i < end; i++)
            singleCost(frames, start, end + 1, i + 1);
        return;
    }
    singleCost(frames, start, end + 1, avg + 1);
    compCostBref(frames, start, avg, avg - start);
    compCostBref(frames, avg + 1, end, end - avg);
}

/* x265_gop_ra (x265.h:771-...): the random-access sub-GOPs of 4 / 8 / 16 pictures in coded order, { POC offset, temporal layer } */
static const signed char s_gopRaLength[3] = { 4, 8, 16 };
static const signed char s_gopRa[3][16][2] = {
    { {4, 0}, {2, 1}, {1, 2}, {3, 2} },
    { {8, 0}, {4, 1}, {2, 2}, {1, 3}, {3, 3}, {6, 2}, {5, 3}, {7, 3} },
    { {16, 0}, {8, 1}, {4, 2}, {2, 3}, {1, 4}, {3, 4}, {6, 3}, {5, 4}, {7, 4}, {12, 2}, {10, 3}, {9, 4}, {11, 4}, {14, 3}, {13, 4}, {15, 4} } };

/* The tail of slicetypeDecide with more than two temporal layers (slicetype.cpp:2061-2325): bframes is 3 / 7 / 15 and
 * b-adapt is off (Encoder::configure, encoder.cpp:3933-3943).  A complete mini-GOP leaves in the coded order of its
 * random-access structure, every frame tagged with its position, structure and temporal layer (the DPB builds the
 * reference picture sets from those); a partial one (scene cut, keyframe, end of stream) is split into the largest
 * smaller structures that fit, each closed by its own P, and a rest of up to three frames coded B-refs first. */
void Lookahead::decideTemporalLayers(Frame** list, Lowres** frames, Frame** fr, int bframes, int brefs, int& maxSearch)
{
    const LookaheadParam& p = m_param;
    const int topGop = p.bEnableTemporalSubLayers - 3;     /* Lookahead::m_gopId, :1097-1114 */
    (void)fr;
    /* close frames list[from .. last] as a sub mini-GOP: types, B-refs, the costs rate control will ask for */
    struct Close
    {
        static void run(Lookahead& la, Frame** list, Lowres** frames, int from, int last, bool bref, int* brefs)
        {
            Lowres& l = list[last]->m_lowres;
            if (!isTypeI(l.sliceType)) l.sliceType = TYPE_P;
            if (last) list[last - 1]->m_lowres.bLastMiniGopBFrame = true;
            l.leadingBframes = last;            /* the index in the whole list, as the reference writes it */
            la.m_lastNonB = &l; la.m_lastNonBFrame = list[last];
            if (bref) la.placeBref(list, from, last, last + 1, brefs);
            if (la.m_param.rc.rateControlMode != 1 /* X265_RC_CQP */)
            {
                const int b = last + 1, p1 = b;
                const int p0 = isTypeI(frames[last + 1]->sliceType) ? b : from;
                la.singleCost(frames, p0, p1, b);
                frames[b]->rcPlanD0 = b - p0; frames[b]->rcPlanD1 = p1 - b;
                if (last) la.compCostBref(frames, from, last, last + 1);
            }
        }
    };
    if (bframes < p.bframes)
    {
        int leftOver = bframes + 1;
        int gopId = topGop - 1;
        int gopLen = gopId >= 0 ? s_gopRaLength[gopId] : 0;
        int listReset = 0;
        while (gopId >= 0 && leftOver > 3 && !m_failed)
        {
            if (leftOver < gopLen)
            {
                gopId--;
                gopLen = gopId >= 0 ? s_gopRaLength[gopId] : 0;
                continue;
            }
            const int newbFrames = listReset + gopLen - 1;
            Close::run(*this, list, frames, listReset, newbFrames, p.bBPyramid && newbFrames, &brefs);
            if (m_failed) return;
            int64_t pts[BFRAME_MAX + 1];
            for (int i = 0; i < gopLen; i++)
            {
                pts[i] = m_inputQueue.front()->m_pts;
                m_inputQueue.pop_front();
                maxSearch--;
            }
            int idx = 0;
            list[newbFrames]->m_reorderedPts = pts[idx++];
            list[newbFrames]->m_gopOffset = 0; list[newbFrames]->m_gopId = gopId; list[newbFrames]->m_gopIdSet = true; list[newbFrames]->m_tempLayer = s_gopRa[gopId][0][1];
            m_outputQueue.push_back(list[newbFrames]);
            for (int j = 1; j < gopLen; j++)
            {
                Frame* f = list[listReset + s_gopRa[gopId][j][0] - 1];
                f->m_gopOffset = j;
                list[bframes]->m_gopId = gopId; list[bframes]->m_gopIdSet = true;     /* sic (:2152): the LAST frame of the list is tagged, not this one */
                f->m_tempLayer = s_gopRa[gopId][j][1];
                f->m_reorderedPts = pts[idx++];
                m_outputQueue.push_back(f);
            }
            listReset += gopLen;
            leftOver -= gopLen;
            gopId--;
            gopLen = gopId >= 0 ? s_gopRaLength[gopId] : 0;
        }
        if (leftOver > 0 && leftOver < 4)
        {
            const int newbFrames = listReset + leftOver - 1;
            Close::run(*this, list, frames, listReset, newbFrames, p.bBPyramid && (newbFrames - listReset) > 1, &brefs);
            if (m_failed) return;
            int64_t pts[BFRAME_MAX + 1];
            for (int i = 0; i < leftOver; i++)
            {
                pts[i] = m_inputQueue.front()->m_pts;
                m_inputQueue.pop_front();
                maxSearch--;
            }
            int idx = 0;
            list[newbFrames]->m_reorderedPts = pts[idx++];
            list[newbFrames]->m_gopOffset = 0; list[newbFrames]->m_gopId = -1; list[newbFrames]->m_gopIdSet = true; list[newbFrames]->m_tempLayer = 0;
            m_outputQueue.push_back(list[newbFrames]);
            if (brefs)
                for (int i = listReset; i < newbFrames; i++)
                    if (list[i]->m_lowres.sliceType == TYPE_BREF)
                    {
                        list[i]->m_reorderedPts = pts[idx++];
                        list[i]->m_gopOffset = 0; list[i]->m_gopId = -1; list[i]->m_gopIdSet = true; list[i]->m_tempLayer = 0;
                        m_outputQueue.push_back(list[i]);
                    }
            for (int i = listReset; i < newbFrames; i++)
                if (list[i]->m_lowres.sliceType != TYPE_BREF)
                {
                    list[i]->m_reorderedPts = pts[idx++];
                    list[i]->m_gopOffset = 0; list[i]->m_gopId = -1; list[i]->m_gopIdSet = true; list[i]->m_tempLayer = 1;
                    m_outputQueue.push_back(list[i]);
                }
        }
        return;
    }
    /* the complete mini-GOP */
    list[bframes - 1]->m_lowres.bLastMiniGopBFrame = true;
    list[bframes]->m_lowres.leadingBframes = bframes;
    m_lastNonB = &list[bframes]->m_lowres; m_lastNonBFrame = list[bframes];
    if (p.bBPyramid && !brefs)
        placeBref(list, 0, bframes, bframes + 1, &brefs);
    if (p.rc.rateControlMode != 1)
    {
        const int b = bframes + 1, p1 = b;
        const int p0 = isTypeI(frames[bframes + 1]->sliceType) ? b : 0;
        singleCost(frames, p0, p1, b);
        frames[b]->rcPlanD0 = b - p0; frames[b]->rcPlanD1 = p1 - b;
        compCostBref(frames, 0, bframes, bframes + 1);
    }
    if (m_failed) return;
    int64_t pts[BFRAME_MAX + 1];
    for (int i = 0; i <= bframes; i++)
    {
        pts[i] = m_inputQueue.front()->m_pts;
        m_inputQueue.pop_front();
        maxSearch--;
    }
    int idx = 0;
    list[bframes]->m_reorderedPts = pts[idx++];
    list[bframes]->m_gopOffset = 0; list[bframes]->m_gopId = topGop; list[bframes]->m_gopIdSet = true; list[bframes]->m_tempLayer = s_gopRa[topGop][0][1];
    m_outputQueue.push_back(list[bframes]);
    for (int j = 1; j <= bframes; j++)
    {
        Frame* f = list[s_gopRa[topGop][j][0] - 1];
        f->m_gopOffset = j; f->m_gopId = topGop; f->m_gopIdSet = true; f->m_tempLayer = s_gopRa[topGop][j][1];
        f->m_reorderedPts = pts[idx++];
        m_outputQueue.push_back(f);
    }
}

void Lookahead::slicetypeDecide()
{
    Lowres* frames[LOOKAHEAD_MAX + BFRAME_MAX + 4];
    Frame*  fr[LOOKAHEAD_MAX + BFRAME_MAX + 4];
    Frame*  list[BFRAME_MAX + 4];
    memset(frames, 0, sizeof(frames)); memset(fr, 0, sizeof(fr)); memset(list, 0, sizeof(list));
    int maxSearch = std::max(1, std::min(m_param.lookaheadDepth, (int)LOOKAHEAD_MAX));

    int j;
    for (j = 0; j < m_param.bframes + 2 && j < (int)m_inputQueue.size(); j++)
        list[j] = m_inputQueue[j];
    frames[0] = m_lastNonB; fr[0] = m_lastNonBFrame;
    std::vector<Frame*> pre;
    for (j = 0; j < maxSearch && j < (int)m_inputQueue.size(); j++)
    {
        fr[j + 1] = m_inputQueue[j];
        frames[j + 1] = &m_inputQueue[j]->m_lowres;
        if (!m_inputQueue[j]->m_lowresInit) pre.push_back(m_inputQueue[j]);
    }
    maxSearch = j;

    for (j = maxSearch; j < m_param.bframes + 2 && j < (int)m_inputQueue.size(); j++)
        if (!m_inputQueue[j]->m_lowresInit) pre.push_back(m_inputQueue[j]);
    const double tStart = nowSec();
    int maxPoc = m_lastNonBFrame ? m_lastNonBFrame->m_poc : -1;
    for (j = 0; j < (int)m_inputQueue.size() && (j < maxSearch || j < m_param.bframes + 2); j++)
        maxPoc = std::max(maxPoc, m_inputQueue[j]->m_poc);
    if (m_param.speculate)
    {
        /* everything the window can ask for goes to the GPU (most of it went long ago, at addPicture) before the
         * host waits for anything */
        /* per-decision mode issues every frame that has arrived, also those beyond the window (asyncDepth), so the
         * GPU works on them while this decision is taken from the results of earlier batches */
        drainPending(m_param.speculate >= 2 ? 0x7fffffff : (size_t)m_param.pendingMax, maxPoc);
        if (m_failed) return;
    }
    {
        /* pixel statistics of the window's frames that nothing above had to wait for yet */
        std::vector<Frame*> still;
        for (size_t i = 0; i < pre.size(); i++)
            if (!pre[i]->m_lowresInit) still.push_back(pre[i]);
        const double t0 = nowSec();
        preLookahead(still);
        m_timers[0] += nowSec() - t0;
    }
    if (m_failed) return;
    verifyWeights(maxPoc);      /* weights assumed for frames this decision reads are settled (and redone) first */
    if (m_failed) return;
    if (m_param.speculate)
    {
        const double t0 = nowSec();
        std::vector<Lowres*> who;
        for (size_t i = 0; i < m_resident.size(); i++)
            if (m_resident[i]->m_poc <= maxPoc && m_resident[i]->m_speculated) who.push_back(&m_resident[i]->m_lowres);
        resolveAlias(who);
        fetchResults(who, maxPoc);
        m_timers[3] += nowSec() - t0;
    }
    if (m_failed) return;
    const double tAnalyse = nowSec();

    const LookaheadParam& p = m_param;
    if (p.bEnableFades)
    {
        /* slicetype.cpp:1861-1906, as is: the frame variances of the mini-GOP candidates in a ring of BFRAME_MAX + 4 entries
         * indexed by POC; a run of non-decreasing variances is a fade-in, and the frame where a fade-in of at least one
         * second stops is marked bIsFadeEnd (coded as a keyframe below; rate control resets on it, ratecontrol.cpp:1416).
         * The walk does not wrap when the candidates straddle the end of the ring (k starts above its end value): kept */
        int j, endIndex = 0;
        const int length = BFRAME_MAX + 4;
        for (j = 0; j < length; j++)
            m_frameVariance[j] = -1;
        for (j = 0; list[j] != NULL; j++)
            m_frameVariance[list[j]->m_poc % length] = list[j]->m_lowres.frameVariance;
        for (int k = list[0]->m_poc % length; k <= list[j - 1]->m_poc % length; k++)
        {
            if (m_frameVariance[k] == -1)
                break;
            if ((k > 0 && m_frameVariance[k] >= m_frameVariance[k - 1]) ||
                (k == 0 && m_frameVariance[k] >= m_frameVariance[length - 1]))
            {
                m_isFadeIn = true;
                if (m_fadeCount == 0 && m_fadeStart == -1)
                {
                    for (int temp = list[0]->m_poc; temp <= list[j - 1]->m_poc; temp++)
                        if (k == temp % length)
                        {
                            m_fadeStart = temp ? temp - 1 : 0;
                            break;
                        }
                }
                m_fadeCount = list[endIndex]->m_poc > m_fadeStart ? list[endIndex]->m_poc - m_fadeStart : 0;
                endIndex++;
            }
            else
            {
                if (m_isFadeIn && m_fadeCount >= p.fpsNum / p.fpsDenom)
                {
                    for (int temp = 0; list[temp] != NULL; temp++)
                        if (list[temp]->m_poc == m_fadeStart + (int)m_fadeCount)
                        {
                            list[temp]->m_lowres.bIsFadeEnd = true;
                            break;
                        }
                }
                m_isFadeIn = false;
                m_fadeCount = 0;
                m_fadeStart = -1;
            }
            if (k == length - 1)
                k = -1;
        }
    }
    if (m_lastNonB && ((p.bFrameAdaptive && p.bframes) || p.rc.cuTree || p.scenecutThreshold || p.bHistBasedSceneCut ||
                       (p.lookaheadDepth && p.rc.vbvBufferSize)))
        slicetypeAnalyse(frames, fr, false);
    if (m_failed) return;

    int bframes, brefs;
    const bool isClosedGopRadl = p.radl && p.keyframeMax != p.keyframeMin;      /* :1933 */
    for (bframes = 0, brefs = 0;; bframes++)
    {
        Lowres& frm = list[bframes]->m_lowres;
        if (frm.sliceTypeReq != TYPE_AUTO && frm.sliceTypeReq != frm.sliceType)
            frm.sliceType = frm.sliceTypeReq;
        if (frm.sliceType == TYPE_BREF && !p.bBPyramid && brefs == p.bBPyramid)
            frm.sliceType = TYPE_B;
        else if (frm.sliceType == TYPE_BREF && p.bBPyramid && brefs && p.maxNumReferences <= (brefs + 3))
            frm.sliceType = TYPE_B;
        if ((!p.bIntraRefresh || frm.frameNum == 0) && frm.frameNum - m_lastKeyframe >= p.keyframeMax &&
            (!m_extendGopBoundary || frm.frameNum - m_lastKeyframe >= p.keyframeMax + p.gopLookahead))
        {
            if (frm.sliceType == TYPE_AUTO || frm.sliceType == TYPE_I)
                frm.sliceType = p.bOpenGOP && m_lastKeyframe >= 0 ? TYPE_I : TYPE_IDR;
            bool warn = frm.sliceType != TYPE_IDR;
            if (warn && p.bOpenGOP) warn &= frm.sliceType != TYPE_I;
            if (warn)
                frm.sliceType = p.bOpenGOP && m_lastKeyframe >= 0 ? TYPE_I : TYPE_IDR;
        }
        if (frm.bIsFadeEnd)         /* :1972 */
            frm.sliceType = p.bOpenGOP && m_lastKeyframe >= 0 ? TYPE_I : TYPE_IDR;
        if (frm.sliceType == TYPE_I && frm.frameNum - m_lastKeyframe >= p.keyframeMin)
        {
            if (p.bOpenGOP) { m_lastKeyframe = frm.frameNum; frm.bKeyframe = true; }
            else frm.sliceType = TYPE_IDR;
        }
        if (frm.sliceType == TYPE_IDR && frm.bScenecut && isClosedGopRadl)      /* --radl, :1995-2000 */
        {
            /* (the reference indexes list[] without checking that the frames exist; a closed-GOP RADL encode whose
             * last scene cut sits in the final frames of the stream crashes there) */
            for (int i = bframes; i < bframes + p.radl && list[i]; i++)
                list[i]->m_lowres.sliceType = TYPE_B;
            if (list[bframes + p.radl])
                list[bframes + p.radl]->m_lowres.sliceType = TYPE_IDR;
        }
        if (frm.sliceType == TYPE_IDR)
        {
            m_lastKeyframe = frm.frameNum;
            frm.bKeyframe = true;
            if (bframes > 0 && !p.radl)
            {
                list[bframes - 1]->m_lowres.sliceType = TYPE_P;
                bframes--;
            }
        }
        if (bframes == p.bframes || !list[bframes + 1])
        {
            if (frm.sliceType == TYPE_AUTO || isTypeB(frm.sliceType))
                frm.sliceType = TYPE_P;
        }
        if (frm.sliceType == TYPE_BREF) brefs++;
        if (frm.sliceType == TYPE_AUTO) frm.sliceType = TYPE_B;
        else if (!isTypeB(frm.sliceType)) break;
    }

    if (p.bEnableTemporalSubLayers > 2)
        decideTemporalLayers(list, frames, fr, bframes, brefs, maxSearch);      /* :2061-2325 */
    else
    {
        if (bframes) list[bframes - 1]->m_lowres.bLastMiniGopBFrame = true;
        list[bframes]->m_lowres.leadingBframes = bframes;
        m_lastNonB = &list[bframes]->m_lowres;
        m_lastNonBFrame = list[bframes];

        if (p.bBPyramid && bframes > 1 && !brefs)
            placeBref(list, 0, bframes, bframes + 1, &brefs);

        /* costs RateControl will ask for (slicetype.cpp:2378-2427) */
        if (p.rc.rateControlMode != 1 /* X265_RC_CQP */)
        {
            int p0, p1, b;
            if (!maxSearch)
                for (int i = 0; i <= bframes; i++) { frames[i + 1] = &list[i]->m_lowres; fr[i + 1] = list[i]; }
            p1 = b = bframes + 1;
            p0 = isTypeI(frames[bframes + 1]->sliceType) ? b : 0;
            singleCost(frames, p0, p1, b);
            frames[b]->rcPlanD0 = b - p0; frames[b]->rcPlanD1 = p1 - b;
            if (p.bEnableTemporalSubLayers > 1 && bframes)
                compCostBref(frames, 0, bframes, bframes + 1);
            else if (bframes)
            {
                p0 = 0;
                bool isp0available = frames[bframes + 1]->sliceType != TYPE_IDR;
                for (b = 1; b <= bframes; b++)
                {
                    if (!isp0available) p0 = b;
                    if (frames[b]->sliceType == TYPE_B)
                        for (p1 = b; frames[p1]->sliceType == TYPE_B; p1++) ;
                    else
                        p1 = bframes + 1;
                    singleCost(frames, p0, p1, b);
                    frames[b]->rcPlanD0 = b - p0; frames[b]->rcPlanD1 = p1 - b;
                    if (frames[b]->sliceType == TYPE_BREF) { p0 = b; isp0available = true; }
                }
            }
        }
        if (m_failed) return;

        /* move the mini-GOP to the output queue in coded order (:2429-2472) */
        int64_t pts[BFRAME_MAX + 1];
        for (int i = 0; i <= bframes; i++)
        {
            pts[i] = m_inputQueue.front()->m_pts;
            m_inputQueue.pop_front();
            maxSearch--;
        }
        int idx = 0;
        list[bframes]->m_reorderedPts = pts[idx++];
        m_outputQueue.push_back(list[bframes]);
        if (brefs)
            for (int i = 0; i < bframes; i++)
                if (list[i]->m_lowres.sliceType == TYPE_BREF)
                {
                    list[i]->m_reorderedPts = pts[idx++];
                    m_outputQueue.push_back(list[i]);
                }
        for (int i = 0; i < bframes; i++)
            if (list[i]->m_lowres.sliceType != TYPE_BREF)
            {
                list[i]->m_reorderedPts = pts[idx++];
                m_outputQueue.push_back(list[i]);
            }
    }
    if (m_failed) return;

    /* keyframe re-analysis for cuTree / VBV (:2475-2504) */
    bool isKeyFrameAnalyse = p.rc.cuTree || (p.rc.vbvBufferSize && p.lookaheadDepth);
    if (isKeyFrameAnalyse && isTypeI(m_lastNonB->sliceType))
    {
        memset(frames, 0, sizeof(frames)); memset(fr, 0, sizeof(fr));
        frames[0] = m_lastNonB; fr[0] = m_lastNonBFrame;
        for (j = 0; j < maxSearch && j < (int)m_inputQueue.size(); j++)
        {
            frames[j + 1] = &m_inputQueue[j]->m_lowres;
            fr[j + 1] = m_inputQueue[j];
        }
        frames[j + 1] = NULL;
        slicetypeAnalyse(frames, fr, true);
    }
    /* The qp offsets of the frames just output are final now (cuTreeFinish ran in the analysis above, or in the keyframe
     * re-analysis): enqueue the cuTree-adjusted cost rate control will ask for (frameCostRecalculate, :3802-3879), so that
     * getEstimatedPictureCost finds it done instead of queueing behind the cuTree passes of later decisions */
    if (p.rc.cuTree && p.rc.rateControlMode != 1 && !(p.shardCount > 1 && m_shardRank > 0))
        for (int i = 0; i <= bframes && !m_failed; i++)
        {
            Lowres& l = list[i]->m_lowres;
            if (l.sliceType == TYPE_B || l.rcPlanD0 < 0) continue;
            const int cs = l.costStore[l.rcPlanD0][l.rcPlanD1];
            if (cs >= 0) check(x265cu_cost_recalc_enqueue(m_ctx, l.slot, cs, 1), "x265cu_cost_recalc_enqueue");
            else l.rcPlanD0 = l.rcPlanD1 = -1;
        }
    m_timers[4] += nowSec() - tAnalyse;
    /* the pictures that were still uploading when this decision started have landed by now: hand them to the GPU
     * before returning to the caller instead of at the next decision */
    if (m_param.speculate == 1 && m_param.shardCount <= 1)
        drainPending((size_t)m_param.pendingMax, -1);
    m_timers[5] += nowSec() - tStart;
    m_timers[6] += 1;
}

/* slicetype.cpp:2603-2919 */
void Lookahead::slicetypeAnalyse(Lowres** frames, Frame** fr, bool bKeyframe)
{
    const LookaheadParam& p = m_param;
    int numFrames, origNumFrames, keyintLimit, framecnt;
    int maxSearch = std::min(p.lookaheadDepth, (int)LOOKAHEAD_MAX);
    int cuCount = m_8x8Blocks;
    int resetStart;
    bool bIsVbvLookahead = p.rc.vbvBufferSize && p.lookaheadDepth;
    (void)fr;

    for (framecnt = 0; framecnt < maxSearch; framecnt++)
    {
        Lowres* fenc = frames[framecnt + 1];
        if (!fenc || fenc->sliceType != TYPE_AUTO)
            break;
    }
    if (!framecnt)
    {
        if (p.rc.cuTree)
            cuTree(frames, 0, bKeyframe);
        return;
    }
    frames[framecnt + 1] = NULL;

    int keyFrameLimit = p.keyframeMax + m_lastKeyframe - frames[0]->frameNum - 1;
    if (p.gopLookahead && keyFrameLimit <= p.bframes + 1)
        keyintLimit = keyFrameLimit + p.gopLookahead;
    else
        keyintLimit = keyFrameLimit;
    origNumFrames = numFrames = p.bIntraRefresh ? framecnt : std::min(framecnt, keyintLimit);
    if (bIsVbvLookahead)
        numFrames = framecnt;
    else if (p.bOpenGOP && numFrames < framecnt)
        numFrames++;
    else if (numFrames == 0)
    {
        frames[1]->sliceType = TYPE_I;
        return;
    }

    if (m_bBatchMotionSearch)
    {
        /* the reference's thread-pool batches (:2668-2736); here they only fix the order of first
         * touch (and with it the context / slicedness of the searches), the work itself was speculated */
        m_inBatch = true;
        for (int b = 2; b < numFrames; b++)
            for (int i = 1; i <= p.bframes + 1; i++)
            {
                int p0 = b - i;
                if (p0 < 0) continue;
                if (frames[b]->mvStore[0][i] >= 0) continue;
                int p1 = b + i;
                if (p1 >= numFrames || frames[b]->mvStore[1][i] >= 0)
                    p1 = b;
                estimateFrameCost(frames, p0, p1, b, false);
            }
        m_bBatchMotionSearch &= p.poolWorkers >= 4;
        if (m_bBatchFrameCosts)
        {
            for (int b = 2; b < numFrames; b++)
                for (int i = 1; i <= p.bframes + 1; i++)
                {
                    if (b < i) continue;
                    if (frames[b]->mvStore[0][i] < 0) continue;
                    int p0 = b - i;
                    for (int jj = 0; jj <= p.bframes; jj++)
                    {
                        int p1 = b + jj;
                        if (p1 >= numFrames) break;
                        if (jj && frames[b]->mvStore[1][jj] < 0) continue;
                        if (frames[b]->costEst[i][jj] >= 0) continue;
                        estimateFrameCost(frames, p0, p1, b, false);
                    }
                }
            m_bBatchFrameCosts &= p.poolWorkers > 12;
        }
        m_inBatch = false;
    }

    int numBFrames = 0, numAnalyzed = numFrames;
    bool isScenecut = p.bHistBasedSceneCut ? histBasedScenecut(frames, 0, 1, origNumFrames)      /* :2742-2745 */
                                           : scenecut(frames, 0, 1, true, origNumFrames);
    if (p.scenecutThreshold && isScenecut)
    {
        frames[1]->sliceType = TYPE_I;
        return;
    }
    if (p.gopLookahead && keyFrameLimit >= 0 && keyFrameLimit <= p.bframes + 1)
    {
        /* a keyframe is due within this mini-GOP: is there a scene cut shortly behind it worth waiting for? (:2753-2771) */
        const bool sceneTransition = m_isSceneTransition;
        m_extendGopBoundary = false;
        for (int i = p.bframes + 1; i < origNumFrames; i += p.bframes + 1)
        {
            scenecut(frames, i, i + 1, true, origNumFrames);
            for (int j = i + 1; j <= std::min(i + p.bframes + 1, origNumFrames); j++)
                if (frames[j]->bScenecut && scenecutInternal(frames, j - 1, j, true))
                {
                    m_extendGopBoundary = true;
                    break;
                }
            if (m_extendGopBoundary)
                break;
        }
        m_isSceneTransition = sceneTransition;
    }
    if (p.bframes)
    {
        if (p.bFrameAdaptive == B_ADAPT_TRELLIS)
        {
            if (numFrames > 1)
            {
                char best_paths[BFRAME_MAX + 1][LOOKAHEAD_MAX + 1];
                memset(best_paths, 0, sizeof(best_paths));
                best_paths[1][0] = 'P';
                int best_path_index = numFrames % (BFRAME_MAX + 1);
                for (int j = 2; j <= numFrames; j++)
                    slicetypePath(frames, j, best_paths);
                numBFrames = (int)strspn(best_paths[best_path_index], "B");
                for (int j = 1; j < numFrames; j++)
                    frames[j]->sliceType = best_paths[best_path_index][j - 1] == 'B' ? TYPE_B : TYPE_P;
            }
            frames[numFrames]->sliceType = TYPE_P;
        }
        else if (p.bFrameAdaptive == B_ADAPT_FAST)
        {
            int64_t cost1p0, cost2p0, cost1b1, cost2p1;
            for (int i = 0; i <= numFrames - 2;)
            {
                cost2p1 = singleCost(frames, i + 0, i + 2, i + 2, true);
                if (frames[i + 2]->intraMbs[2] > cuCount / 2)
                {
                    frames[i + 1]->sliceType = TYPE_P;
                    frames[i + 2]->sliceType = TYPE_P;
                    i += 2;
                    continue;
                }
                cost1b1 = singleCost(frames, i + 0, i + 2, i + 1);
                cost1p0 = singleCost(frames, i + 0, i + 1, i + 1);
                cost2p0 = singleCost(frames, i + 1, i + 2, i + 2);
                if (cost1p0 + cost2p0 < cost1b1 + cost2p1)
                {
                    frames[i + 1]->sliceType = TYPE_P;
                    i += 1;
                    continue;
                }
                frames[i + 1]->sliceType = TYPE_B;
                int j;
                for (j = i + 2; j <= std::min(i + p.bframes, numFrames - 1); j++)
                {
                    int64_t pthresh = std::max(300 - (50 - p.bFrameBias) * (j - i - 1), 300 / 10);
                    int64_t pcost = singleCost(frames, i + 0, j + 1, j + 1, true);
                    if (pcost > pthresh * cuCount || frames[j + 1]->intraMbs[j - i + 1] > cuCount / 3)
                        break;
                    frames[j]->sliceType = TYPE_B;
                }
                frames[j]->sliceType = TYPE_P;
                i = j;
            }
            frames[numFrames]->sliceType = TYPE_P;
            numBFrames = 0;
            while (numBFrames < numFrames && frames[numBFrames + 1]->sliceType == TYPE_B)
                numBFrames++;
        }
        else
        {
            numBFrames = std::min(numFrames - 1, p.bframes);
            for (int j = 1; j < numFrames; j++)
                frames[j]->sliceType = (j % (numBFrames + 1)) ? TYPE_B : TYPE_P;
            frames[numFrames]->sliceType = TYPE_P;
        }
        /* scenecut check on the first mini-GOP (:2868-2880) */
        for (int j = 1; j < numBFrames + 1; j++)
            if (scenecut(frames, j, j + 1, false, origNumFrames))
            {
                frames[j]->sliceType = TYPE_P;
                numAnalyzed = j;
                break;
            }
        resetStart = bKeyframe ? 1 : std::min(numBFrames + 2, numAnalyzed + 1);
    }
    else
    {
        for (int j = 1; j <= numFrames; j++)
            frames[j]->sliceType = TYPE_P;
        resetStart = bKeyframe ? 1 : 2;
    }

    if (p.rc.cuTree)
        cuTree(frames, std::min(numFrames, p.keyframeMax), bKeyframe);

    if (p.gopLookahead && keyFrameLimit >= 0 && keyFrameLimit <= p.bframes + 1 && !m_extendGopBoundary)   /* :2896-2897 */
        keyintLimit = keyFrameLimit;

    if (!p.bIntraRefresh)
        for (int j = keyintLimit + 1; j <= numFrames; j += p.keyframeMax)
        {
            frames[j]->sliceType = TYPE_I;
            resetStart = std::min(resetStart, j + 1);
        }

    if (bIsVbvLookahead)
        vbvLookahead(frames, numFrames, bKeyframe);
    int maxp1 = std::min(p.bframes + 1, origNumFrames);
    for (int j = resetStart; j <= numFrames; j++)
    {
        frames[j]->sliceType = TYPE_AUTO;
        if (j <= maxp1 && frames[j]->bScenecut && m_isSceneTransition)
            m_isSceneTransition = false;
    }
}

/* slicetype.cpp:2921-3014 */
bool Lookahead::scenecut(Lowres** frames, int p0, int p1, bool bRealScenecut, int numFrames)
{
    if (bRealScenecut && m_param.bframes)
    {
        int origmaxp1 = p0 + 1 + m_param.bframes;
        int maxp1 = std::min(origmaxp1, numFrames);
        bool fluctuate = false, noScenecuts = false;
        int64_t avgSatdCost = 0;
        if (frames[p0]->costEst[p1 - p0][0] > -1)
            avgSatdCost = frames[p0]->costEst[p1 - p0][0];
        int cnt = 1;
        for (int cp1 = p1; cp1 <= maxp1; cp1++)
        {
            if (!scenecutInternal(frames, p0, cp1, false))
            {
                for (int i = cp1; i > p0; i--)
                {
                    frames[i]->bScenecut = false;
                    noScenecuts = false;
                }
            }
            else if (scenecutInternal(frames, cp1 - 1, cp1, false))
            {
                frames[cp1]->bScenecut = true;
                noScenecuts = true;
            }
            avgSatdCost += frames[cp1]->costEst[cp1 - p0][0];
            cnt++;
        }
        if (noScenecuts)
        {
            fluctuate = false;
            avgSatdCost /= cnt;
            for (int i = p1; i <= maxp1; i++)
            {
                int64_t curCost = frames[i]->costEst[i - p0][0];
                int64_t prevCost = frames[i - 1]->costEst[i - 1 - p0][0];
                if (fabs((double)(curCost - avgSatdCost)) > 0.1 * avgSatdCost ||
                    fabs((double)(curCost - prevCost)) > 0.1 * prevCost)
                {
                    fluctuate = true;
                    if (!m_isSceneTransition && frames[i]->bScenecut)
                    {
                        m_isSceneTransition = true;
                        for (int j = i + 1; j <= maxp1; j++)
                            frames[j]->bScenecut = false;
                        break;
                    }
                }
                frames[i]->bScenecut = false;
            }
        }
        if (!fluctuate && !noScenecuts)
            m_isSceneTransition = false;
    }
    if (m_param.csvLogLevel >= 2)      /* :2998-3003 */
    {
        int64_t icost = frames[p1]->costEst[0][0];
        int64_t pcost = frames[p1]->costEst[p1 - p0][0];
        frames[p1]->ipCostRatio = (double)icost / pcost;
    }
    if (!frames[p1]->bScenecut)
        return false;
    return scenecutInternal(frames, p0, p1, bRealScenecut);
}

/* slicetype.cpp:3016-3055 */
/* slicetype.cpp:3057-3188, as is -- including the segment size that keeps growing by the last segment's remainder from one
 * segment to the next (it only feeds the thresholds) */
bool Lookahead::detectHistBasedSceneChange(Lowres** frames, int p0, int p1, int p2)
{
    enum { PICTURE_DIFF_VARIANCE_TH = 390, PICTURE_VARIANCE_TH = 1500, LOW_VAR_SCENE_CHANGE_TH = 2250, HIGH_VAR_SCENE_CHANGE_TH = 3500,
           PICTURE_DIFF_VARIANCE_CHROMA_TH = 10, PICTURE_VARIANCE_CHROMA_TH = 20, LOW_VAR_SCENE_CHANGE_CHROMA_TH = 2250 / 4,
           HIGH_VAR_SCENE_CHANGE_CHROMA_TH = 3500 / 4, FADE_TH = 4, INTENSITY_CHANGE_TH = 4 };       /* slicetype.h:49-61 */
    const double FLASH_TH = 1.5;
    Lowres* previousFrame = frames[p0];
    Lowres* currentFrame = frames[p1];
    Lowres* futureFrame = frames[p2];
    currentFrame->bHistScenecutAnalyzed = true;
    if (!previousFrame->hist || !currentFrame->hist || !futureFrame->hist) { fail("hist-scenecut statistics missing"); return false; }
    const x265cu_hist_stats &prev = *previousFrame->hist, &cur = *currentFrame->hist, &fut = *futureFrame->hist;

    uint8_t absIntDiffFuturePast = 0, absIntDiffFuturePresent = 0, absIntDiffPresentPast = 0;
    uint32_t abruptChangeCount = 0, sceneChangeCount = 0;
    uint32_t segmentWidth = (uint32_t)m_param.sourceWidth / 4, segmentHeight = (uint32_t)m_param.sourceHeight / 4;
#define LA_NUM64(w, h) (((w) * (h)) >> (6 << 1))     /* NUM64x64INPIC, MAX_LOG2_CU_SIZE = 6 */
    for (uint32_t wi = 0; wi < 4; wi++)
    {
        for (uint32_t hi = 0; hi < 4; hi++)
        {
            bool isAbruptChange = false, isSceneChange = false;
            uint32_t accHistDiff = 0, accHistDiffCb = 0, accHistDiffCr = 0;
            uint32_t segmentWidthOffset = wi == 3 ? (uint32_t)m_param.sourceWidth - 4 * segmentWidth : 0;
            uint32_t segmentHeightOffset = hi == 3 ? (uint32_t)m_param.sourceHeight - 4 * segmentHeight : 0;
            segmentWidth += segmentWidthOffset;
            segmentHeight += segmentHeightOffset;

            const int64_t dv = std::abs((int64_t)cur.pic_avg_variance[0] - (int64_t)prev.pic_avg_variance[0]);
            uint32_t segmentThreshHold = (dv > PICTURE_DIFF_VARIANCE_TH &&
                                          (cur.pic_avg_variance[0] > PICTURE_VARIANCE_TH || prev.pic_avg_variance[0] > PICTURE_VARIANCE_TH))
                                         ? HIGH_VAR_SCENE_CHANGE_TH * LA_NUM64(segmentWidth, segmentHeight)
                                         : LOW_VAR_SCENE_CHANGE_TH * LA_NUM64(segmentWidth, segmentHeight);
            const int64_t dvb = std::abs((int64_t)cur.pic_avg_variance[1] - (int64_t)prev.pic_avg_variance[1]);
            uint32_t segmentThreshHoldCb = (dvb > PICTURE_DIFF_VARIANCE_CHROMA_TH &&
                                            (cur.pic_avg_variance[1] > PICTURE_VARIANCE_CHROMA_TH || prev.pic_avg_variance[1] > PICTURE_VARIANCE_CHROMA_TH))
                                           ? HIGH_VAR_SCENE_CHANGE_CHROMA_TH * LA_NUM64(segmentWidth, segmentHeight)
                                           : LOW_VAR_SCENE_CHANGE_CHROMA_TH * LA_NUM64(segmentWidth, segmentHeight);
            const int64_t dvr = std::abs((int64_t)cur.pic_avg_variance[2] - (int64_t)prev.pic_avg_variance[2]);
            uint32_t segmentThreshHoldCr = (dvr > PICTURE_DIFF_VARIANCE_CHROMA_TH &&
                                            (cur.pic_avg_variance[2] > PICTURE_VARIANCE_CHROMA_TH || prev.pic_avg_variance[2] > PICTURE_VARIANCE_CHROMA_TH))
                                           ? HIGH_VAR_SCENE_CHANGE_CHROMA_TH * LA_NUM64(segmentWidth, segmentHeight)
                                           : LOW_VAR_SCENE_CHANGE_CHROMA_TH * LA_NUM64(segmentWidth, segmentHeight);

            for (uint32_t bin = 0; bin < 256; ++bin)
            {
                accHistDiff += (uint32_t)std::abs((int32_t)cur.histogram[wi][hi][0][bin] - (int32_t)prev.histogram[wi][hi][0][bin]);
                accHistDiffCb += (uint32_t)std::abs((int32_t)cur.histogram[wi][hi][1][bin] - (int32_t)prev.histogram[wi][hi][1][bin]);
                accHistDiffCr += (uint32_t)std::abs((int32_t)cur.histogram[wi][hi][2][bin] - (int32_t)prev.histogram[wi][hi][2][bin]);
            }
            if (m_resetRunningAvg)
            {
                m_accHistDiffRunningAvg[wi][hi] = accHistDiff;
                m_accHistDiffRunningAvgCb[wi][hi] = accHistDiffCb;
                m_accHistDiffRunningAvgCr[wi][hi] = accHistDiffCr;
            }
            uint32_t accHistDiffError = (uint32_t)std::abs((int32_t)m_accHistDiffRunningAvg[wi][hi] - (int32_t)accHistDiff);
            uint32_t accHistDiffErrorCb = (uint32_t)std::abs((int32_t)m_accHistDiffRunningAvgCb[wi][hi] - (int32_t)accHistDiffCb);
            uint32_t accHistDiffErrorCr = (uint32_t)std::abs((int32_t)m_accHistDiffRunningAvgCr[wi][hi] - (int32_t)accHistDiffCr);

            if ((accHistDiffError > segmentThreshHold && accHistDiff >= accHistDiffError) ||
                (accHistDiffErrorCb > segmentThreshHoldCb && accHistDiffCb >= accHistDiffErrorCb) ||
                (accHistDiffErrorCr > segmentThreshHoldCr && accHistDiffCr >= accHistDiffErrorCr))
                isAbruptChange = true;

            if (isAbruptChange)
            {
                absIntDiffFuturePast = (uint8_t)std::abs((int16_t)fut.avg_intensity_seg[wi][hi][0] - (int16_t)prev.avg_intensity_seg[wi][hi][0]);
                absIntDiffFuturePresent = (uint8_t)std::abs((int16_t)fut.avg_intensity_seg[wi][hi][0] - (int16_t)cur.avg_intensity_seg[wi][hi][0]);
                absIntDiffPresentPast = (uint8_t)std::abs((int16_t)cur.avg_intensity_seg[wi][hi][0] - (int16_t)prev.avg_intensity_seg[wi][hi][0]);
                if (absIntDiffFuturePresent >= FLASH_TH * absIntDiffFuturePast && absIntDiffPresentPast >= FLASH_TH * absIntDiffFuturePast)
                    ;   /* flash */
                else if (absIntDiffFuturePresent < FADE_TH && absIntDiffPresentPast < FADE_TH)
                    ;   /* fade */
                else if (std::abs(absIntDiffFuturePresent - absIntDiffPresentPast) < INTENSITY_CHANGE_TH &&
                         absIntDiffFuturePresent + absIntDiffPresentPast >= absIntDiffFuturePast)
                    ;   /* intensity change */
                else
                    isSceneChange = true;
            }
            else
                m_accHistDiffRunningAvg[wi][hi] = (3 * m_accHistDiffRunningAvg[wi][hi] + accHistDiff) / 4;

            abruptChangeCount += isAbruptChange;
            sceneChangeCount += isSceneChange;
        }
    }
#undef LA_NUM64
    m_resetRunningAvg = abruptChangeCount >= m_segmentCountThreshold;
    return sceneChangeCount >= m_segmentCountThreshold;
}

/* slicetype.cpp:3190-3216 */
bool Lookahead::histBasedScenecut(Lowres** frames, int p0, int p1, int numFrames)
{
    /* Only do analysis during a normal scenecut check. */
    if (m_param.bframes)
    {
        int origmaxp1 = p0 + 1;
        /* Look ahead to avoid coding short flashes as scenecuts. */
        origmaxp1 += m_param.bframes;
        int maxp1 = std::min(origmaxp1, numFrames);
        for (int cp1 = p0; cp1 < maxp1; cp1++)
        {
            if (frames[cp1 + 1]->bHistScenecutAnalyzed == true)
                continue;
            if (frames[cp1 + 2] != NULL && detectHistBasedSceneChange(frames, cp1, cp1 + 1, cp1 + 2))
                frames[cp1 + 1]->bScenecut = true;
        }
    }
    return frames[p1]->bScenecut;
}

bool Lookahead::scenecutInternal(Lowres** frames, int p0, int p1, bool bRealScenecut)
{
    Lowres* frame = frames[p1];
    singleCost(frames, p0, p1, p1);
    int64_t icost = frame->costEst[0][0];
    int64_t pcost = frame->costEst[p1 - p0][0];
    int gopSize = (frame->frameNum - m_lastKeyframe) % m_param.keyframeMax;
    float threshMax = (float)(m_param.scenecutThreshold / 100.0);
    float threshMin = (float)(threshMax * 0.25);
    double bias = m_param.scenecutBias;
    if (bRealScenecut)
    {
        if (m_param.keyframeMin == m_param.keyframeMax)
            threshMin = threshMax;
        if (gopSize <= m_param.keyframeMin / 4 || m_param.bIntraRefresh)
            bias = threshMin / 4;
        else if (gopSize <= m_param.keyframeMin)
            bias = threshMin * gopSize / m_param.keyframeMin;
        else
            bias = threshMin + (threshMax - threshMin) * (gopSize - m_param.keyframeMin) /
                   (m_param.keyframeMax - m_param.keyframeMin);
    }
    return pcost >= (1.0 - bias) * icost;
}

/* slicetype.cpp:3218-3245 */
void Lookahead::slicetypePath(Lowres** frames, int length, char (*best_paths)[LOOKAHEAD_MAX + 1])
{
    char paths[2][LOOKAHEAD_MAX + 1];
    int num_paths = std::min(m_param.bframes + 1, length);
    int64_t best_cost = 1LL << 62;
    int idx = 0;
    for (int path = 0; path < num_paths; path++)
    {
        int len = length - (path + 1);
        memcpy(paths[idx], best_paths[len % (BFRAME_MAX + 1)], len);
        memset(paths[idx] + len, 'B', path);
        strcpy(paths[idx] + len + path, "P");
        int64_t cost = slicetypePathCost(frames, paths[idx], best_cost);
        if (cost < best_cost)
        {
            best_cost = cost;
            idx ^= 1;
        }
    }
    memcpy(best_paths[length % (BFRAME_MAX + 1)], paths[idx ^ 1], length);
}

/* slicetype.cpp:3268-3313 */
int64_t Lookahead::slicetypePathCost(Lowres** frames, char* path, int64_t threshold)
{
    int64_t cost = 0;
    int loc = 1, cur_p = 0;
    path--;
    while (path[loc])
    {
        int next_p = loc;
        while (path[next_p] != 'P')
            next_p++;
        cost += singleCost(frames, cur_p, next_p, next_p);
        if (cost > threshold)
            break;
        if (m_param.bBPyramid && next_p - cur_p > 2)
        {
            int middle = cur_p + (next_p - cur_p) / 2;
            cost += singleCost(frames, cur_p, next_p, middle);
            for (int next_b = loc; next_b < middle && cost < threshold; next_b++)
                cost += singleCost(frames, cur_p, middle, next_b);
            for (int next_b = middle + 1; next_b < next_p && cost < threshold; next_b++)
                cost += singleCost(frames, middle, next_p, next_b);
        }
        else
        {
            for (int next_b = loc; next_b < next_p && cost < threshold; next_b++)
                cost += singleCost(frames, cur_p, next_p, next_b);
        }
        loc = next_p + 1;
        cur_p = next_p;
    }
    return cost;
}

/* -------------------------------------------------------------------------------------------
 * cuTree (slicetype.cpp:3399-3500) -- the chain walk is host logic, each step is a GPU launch
 * ------------------------------------------------------------------------------------------- */

void Lookahead::cuTree(Lowres** frames, int numframes, bool bIntra)
{
    const LookaheadParam& p = m_param;
    if (m_param.shardCount > 1 && m_shardRank > 0)
        return;     /* not the decision rank: its qp offsets are never read */
    int idx = !bIntra;
    int lastnonb, curnonb = 1;
    int bframes = 0;

    double totalDuration = 0.0;
    for (int j = 0; j <= numframes; j++)
        totalDuration += (double)p.fpsDenom / p.fpsNum;
    double averageDuration = totalDuration / (numframes + 1);

    int i = numframes;
    while (i > 0 && frames[i]->sliceType == TYPE_B)
        i--;
    lastnonb = i;

    if (!p.lookaheadDepth)
    {
        fail("cuTree with rc-lookahead 0 is not supported");
        return;
    }
    if (lastnonb < idx)
        return;
    check(x265cu_cutree_reset(m_ctx, frames[lastnonb]->slot), "x265cu_cutree_reset");

    while (i-- > idx)
    {
        curnonb = i;
        while (frames[curnonb]->sliceType == TYPE_B && curnonb > 0)
            curnonb--;
        if (curnonb < idx)
            break;
        singleCost(frames, curnonb, lastnonb, lastnonb);
        check(x265cu_cutree_reset(m_ctx, frames[curnonb]->slot), "x265cu_cutree_reset");
        bframes = lastnonb - curnonb - 1;
        if (p.bBPyramid && bframes > 1)
        {
            int middle = (bframes + 1) / 2 + curnonb;
            singleCost(frames, curnonb, lastnonb, middle);
            check(x265cu_cutree_reset(m_ctx, frames[middle]->slot), "x265cu_cutree_reset");
            while (i > curnonb)
            {
                int p0 = i > middle ? middle : curnonb;
                int p1 = i < middle ? middle : lastnonb;
                if (i != middle)
                {
                    singleCost(frames, p0, p1, i);
                    estimateCUPropagate(frames, averageDuration, p0, p1, i, 0);
                }
                i--;
            }
            estimateCUPropagate(frames, averageDuration, curnonb, lastnonb, middle, 1);
        }
        else
        {
            while (i > curnonb)
            {
                singleCost(frames, curnonb, lastnonb, i);
                estimateCUPropagate(frames, averageDuration, curnonb, lastnonb, i, 0);
                i--;
            }
        }
        estimateCUPropagate(frames, averageDuration, curnonb, lastnonb, lastnonb, 1);
        lastnonb = curnonb;
        if (m_failed) return;
    }

    cuTreeFinish(frames[lastnonb], averageDuration, lastnonb);
    if (p.bBPyramid && bframes > 1 && !p.rc.vbvBufferSize)
        cuTreeFinish(frames[lastnonb + (bframes + 1) / 2], averageDuration, 0);
}

/* slicetype.cpp:3502-3608 */
void Lookahead::estimateCUPropagate(Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced)
{
    const LookaheadParam& p = m_param;
    int32_t distScaleFactor = (((b - p0) << 8) + ((p1 - p0) >> 1)) / (p1 - p0);
    int32_t bipredWeight = p.bEnableWeightedBiPred ? 64 - (distScaleFactor >> 2) : 32;
    double fpsFactor = clipDuration((double)p.fpsDenom / p.fpsNum) / clipDuration(averageDuration);
    Lowres* fb = frames[b];
    int cs = fb->costStore[b - p0][p1 - b];
    int l0 = fb->mvStore[0][b - p0];
    int l1 = p1 > b ? fb->mvStore[1][p1 - b] : -1;
    if (cs < 0 || l0 < 0) { fail("cuTree propagate on an estimate that was never computed"); return; }
    check(x265cu_cutree_propagate(m_ctx, fb->slot, frames[p0]->slot, frames[p1]->slot, cs, l0, l1,
                                  referenced, bipredWeight, fpsFactor), "x265cu_cutree_propagate");
    if (p.rc.vbvBufferSize && p.lookaheadDepth && referenced)
        cuTreeFinish(frames[b], averageDuration, b == p1 ? b - p0 : 0);
}

/* slicetype.cpp:3750-3798 (non-hevc-aq, qg-size > 8) */
void Lookahead::cuTreeFinish(Lowres* frame, double averageDuration, int ref0Distance)
{
    const LookaheadParam& p = m_param;
    int fpsFactor = (int)(clipDuration(averageDuration) / clipDuration((double)p.fpsDenom / p.fpsNum) * 256);
    double weightdelta = 0.0;
    if (ref0Distance && frame->weightedCostDelta[ref0Distance - 1] > 0)
        weightdelta = (1.0 - frame->weightedCostDelta[ref0Distance - 1]);
    check(x265cu_cutree_finish(m_ctx, frame->slot, fpsFactor, weightdelta, m_cuTreeStrength), "x265cu_cutree_finish");
}

/* slicetype.cpp:3802-3879 */
int64_t Lookahead::frameCostRecalculate(Lowres** frames, int p0, int p1, int b)
{
    if (frames[b]->sliceType == TYPE_B)
        return frames[b]->costEstAq[b - p0][p1 - b];
    if (m_param.shardCount > 1 && m_shardRank > 0)
        return 0;   /* planned costs are only read on the decision rank */
    int64_t score = 0;
    int cs = frames[b]->costStore[b - p0][p1 - b];
    if (cs < 0) { fail("frameCostRecalculate on an estimate that was never computed"); return 0; }
    check(x265cu_cost_recalc(m_ctx, frames[b]->slot, cs, 1, &score, NULL), "x265cu_cost_recalc");
    return score;
}

/* slicetype.cpp:2587-2601 */
int64_t Lookahead::vbvFrameCost(Lowres** frames, int p0, int p1, int b)
{
    int64_t cost = singleCost(frames, p0, p1, b);
    if (m_param.rc.aqMode)
    {
        if (m_param.rc.cuTree)
            return frameCostRecalculate(frames, p0, p1, b);
        else
            return frames[b]->costEstAq[b - p0][p1 - b];
    }
    return cost;
}

/* slicetype.cpp:2510-2585 */
void Lookahead::vbvLookahead(Lowres** frames, int numFrames, int keyframe)
{
    int prevNonB = 0, curNonB = 1, idx = 0;
    while (curNonB < numFrames && isTypeB(frames[curNonB]->sliceType))
        curNonB++;
    int nextNonB = keyframe ? prevNonB : curNonB;
    int nextB = prevNonB + 1;
    int nextBRef = 0, curBRef = 0;
    if (m_param.bBPyramid && curNonB - prevNonB > 1)
        curBRef = (prevNonB + curNonB + 1) / 2;
    int miniGopEnd = keyframe ? prevNonB : curNonB;
    while (curNonB <= numFrames)
    {
        if (nextNonB != curNonB)
        {
            int p0 = isTypeI(frames[curNonB]->sliceType) ? curNonB : prevNonB;
            frames[nextNonB]->plannedSatd[idx] = vbvFrameCost(frames, p0, curNonB, curNonB);
            frames[nextNonB]->plannedType[idx] = frames[curNonB]->sliceType;
            if (curNonB > miniGopEnd)
                for (int j = nextB; j < miniGopEnd; j++)
                {
                    frames[j]->plannedSatd[frames[j]->indB] = frames[nextNonB]->plannedSatd[idx];
                    frames[j]->plannedType[frames[j]->indB++] = frames[nextNonB]->plannedType[idx];
                }
            idx++;
        }
        if (m_param.bBPyramid && curNonB - prevNonB > 1)
            nextBRef = (prevNonB + curNonB + 1) / 2;
        for (int i = prevNonB + 1; i < curNonB; i++, idx++)
        {
            int64_t satdCost = 0;
            int type = TYPE_B;
            if (nextBRef)
            {
                if (i == nextBRef)
                {
                    satdCost = vbvFrameCost(frames, prevNonB, curNonB, nextBRef);
                    type = TYPE_BREF;
                }
                else if (i < nextBRef)
                    satdCost = vbvFrameCost(frames, prevNonB, nextBRef, i);
                else
                    satdCost = vbvFrameCost(frames, nextBRef, curNonB, i);
            }
            else
                satdCost = vbvFrameCost(frames, prevNonB, curNonB, i);
            frames[nextNonB]->plannedSatd[idx] = satdCost;
            frames[nextNonB]->plannedType[idx] = type;
            for (int j = nextB; j < miniGopEnd; j++)
            {
                if (curBRef && curBRef == i)
                    break;
                if (j >= i && j != nextBRef)
                    continue;
                frames[j]->plannedSatd[frames[j]->indB] = satdCost;
                frames[j]->plannedType[frames[j]->indB++] = type;
            }
        }
        prevNonB = curNonB;
        curNonB++;
        while (curNonB <= numFrames && isTypeB(frames[curNonB]->sliceType))
            curNonB++;
    }
    frames[nextNonB]->plannedType[idx] = TYPE_AUTO;
}

/* slicetype.cpp:1327-1386.  The reference derives p0 / p1 from the slice's reference lists; the caller passes the
 * frames themselves (NULL = none) or their POC distances. */
void Lookahead::getEstimatedPictureCost(Frame* cur, Frame* ref0, Frame* ref1)
{
    getEstimatedPictureCost(cur, ref0 ? cur->m_poc - ref0->m_poc : 0, ref1 ? ref1->m_poc - cur->m_poc : 0);
}

void Lookahead::getEstimatedPictureCost(Frame* cur, int dist0, int dist1)
{
    int d0, d1;
    const int type = cur->m_lowres.sliceType;
    if (isTypeI(type))
        d0 = d1 = 0;
    else if (type == TYPE_P)
    {
        if (dist0 <= 0) return;
        d0 = dist0; d1 = 0;
    }
    else
    {
        if (dist1 <= 0) return;
        /* a B slice without a list-0 reference (a RADL leading picture) is estimated against itself, :1360-1366 */
        d0 = dist0 > 0 ? dist0 : 0;
        d1 = dist1;
    }
    if (d0 >= m_geom.nb || d1 >= m_geom.nb) { fail("getEstimatedPictureCost: references outside the (bframes+2) window"); return; }
    const double t0 = nowSec();
    Lowres& l = cur->m_lowres;
    l.rcD0 = d0; l.rcD1 = d1;
    if (m_param.rc.cuTree)
    {
        /* frameCostRecalculate (:3802-3879) only reads frames[b] */
        if (l.sliceType == TYPE_B)
            l.satdCost = l.costEstAq[d0][d1];
        else
        {
            int64_t score = 0;
            const int cs = l.costStore[d0][d1];
            if (cs < 0)
            {
                /* rate control names references the lookahead never estimated this frame against.  The reference re-sums whatever its
                 * lowresCosts array of that pair holds (never-written memory); here the call reports it and the stream goes on */
                snprintf(m_error, sizeof(m_error), "getEstimatedPictureCost(%d, %d) of poc %d: that estimate was never computed", d0, d1, cur->m_poc);
                l.satdCost = l.costEst[d0][d1];
                l.rcD0 = l.rcD1 = -1;
                m_timers[8] += nowSec() - t0;
                return;
            }
            if (l.rcPlanD0 == d0 && l.rcPlanD1 == d1)
            {
                check(x265cu_cost_recalc_get(m_ctx, l.slot, cs, &score, NULL), "x265cu_cost_recalc_get");
                l.rcPlanD0 = l.rcPlanD1 = -1;
            }
            else
                check(x265cu_cost_recalc(m_ctx, l.slot, cs, 1, &score, NULL), "x265cu_cost_recalc");
            l.satdCost = score;
        }
    }
    else if (m_param.rc.aqMode)
        l.satdCost = l.costEstAq[d0][d1];
    else
        l.satdCost = l.costEst[d0][d1];
    m_timers[8] += nowSec() - t0;
}

/* The VBV half of getEstimatedPictureCost (slicetype.cpp:1387-1436) for the estimate the last getEstimatedPictureCost of
 * this frame named: per CTU row the sums the reference adds to m_rowStat[].satdForVbv / intraSatdForVbv, and the scaled
 * lowresCostForRc / intraCost arrays it leaves in the Lowres (any output may be NULL). */
bool Lookahead::getVbvRowCosts(Frame* cur, int pirStartCol, int pirEndCol, uint32_t* satdForVbv, uint32_t* intraSatdForVbv,
                               uint16_t* lowresCostForRc, int32_t* intraCostScaled)
{
    const Lowres& l = cur->m_lowres;
    if (l.rcD0 < 0) { snprintf(m_error, sizeof(m_error), "getVbvRowCosts of poc %d without a usable getEstimatedPictureCost before it", cur->m_poc); return false; }
    const int cs = l.costStore[l.rcD0][l.rcD1];
    if (cs < 0)
    {
        /* as above: not fatal.  The row sums stay what the caller initialised them to */
        snprintf(m_error, sizeof(m_error), "getVbvRowCosts(%d, %d) of poc %d: that estimate was never computed", l.rcD0, l.rcD1, cur->m_poc);
        return false;
    }
    const int scale = m_param.maxCUSize / 16;
    /* :1408-1410: B frames and runs without cuTree read qpAqOffset */
    int qpSource = 0;
    if (m_param.rc.aqMode)
        qpSource = (l.sliceType == TYPE_B || !m_param.rc.cuTree) ? 1 : 2;
    const bool pir = m_param.bIntraRefresh && l.sliceType == TYPE_P && pirStartCol >= 0;
    return check(x265cu_vbv_row_costs(m_ctx, l.slot, cs, qpSource, scale, pir ? pirStartCol : -1, pir ? pirEndCol : -1,
                                      vbvRows(), satdForVbv, intraSatdForVbv, lowresCostForRc, intraCostScaled), "x265cu_vbv_row_costs");
}

/* -------------------------------------------------------------------------------------------
 * host mirrors
 * ------------------------------------------------------------------------------------------- */

bool Lookahead::fetchMvs(Frame* f, int list, int dist, int32_t* mvXY, int32_t* mvCosts)
{
    int store = f->m_lowres.mvStore[list][dist];
    if (store < 0)
    {
        if (mvXY) mvXY[0] = 0x7FFF;      /* the reference's "not searched" sentinel, lowres.cpp:355-359 */
        return false;
    }
    return check(x265cu_fetch_mvs(m_ctx, f->m_lowres.slot, store, mvXY, mvCosts), "x265cu_fetch_mvs");
}

bool Lookahead::fetchHmeMvs(Frame* f, int list, int dist, int32_t* mvXY, int32_t* mvCosts)
{
    const int store = f->m_lowres.mvStore[list][dist];
    if (store < 0 || !m_param.bEnableHME) return false;
    /* a sharded stream exchanges the lowres results only: the level-0 vectors stay on the rank that searched the frame */
    if (m_param.shardCount > 1 && f->m_poc % m_param.shardCount != m_shardRank) return false;
    return check(x265cu_fetch_hme_mvs(m_ctx, f->m_lowres.slot, store, mvXY, mvCosts), "x265cu_fetch_hme_mvs");
}

bool Lookahead::fetchCosts(Frame* f, int d0, int d1, uint16_t* lowresCosts, int32_t* rowSatds)
{
    int cs = f->m_lowres.costStore[d0][d1];
    if (cs < 0)
    {
        if (rowSatds) rowSatds[0] = -1;
        return false;
    }
    if (d0 == 0 && d1 == 0)
    {
        x265cu_frame_out o;
        memset(&o, 0, sizeof(o));
        o.lowres_costs00 = lowresCosts; o.row_satds00 = rowSatds;
        return check(x265cu_fetch_frame(m_ctx, f->m_lowres.slot, &o), "x265cu_fetch_frame");
    }
    return check(x265cu_fetch_costs(m_ctx, f->m_lowres.slot, cs, lowresCosts, rowSatds), "x265cu_fetch_costs");
}

bool Lookahead::mirror(Frame* f, const x265cu_mirror_request* req, int64_t* ticket)
{
    const double t0 = nowSec();
    const bool ok = check(x265cu_mirror_enqueue(m_ctx, f->m_lowres.slot, req, ticket), "x265cu_mirror_enqueue");
    m_timers[9] += nowSec() - t0;
    return ok;
}

bool Lookahead::fetchFrame(Frame* f, const x265cu_frame_out* out)
{
    const double t0 = nowSec();
    const bool ok = check(x265cu_fetch_frame(m_ctx, f->m_lowres.slot, out), "x265cu_fetch_frame");
    m_timers[9] += nowSec() - t0;
    return ok;
}

} // namespace x265cu
