/* lookahead.h -- host side of the B200 lookahead, above the C ABI in include/x265cu.h.
 *
 * Mirrors the reference's `class Lookahead` (source/encoder/slicetype.h:151-259): same public
 * method names, argument meaning and error behaviour (bool returns, no exceptions), and the
 * same `Lowres` output fields (source/common/lowres.h:171-263) that RateControl / Analysis /
 * weightPrediction consume.  All decision logic (queues, scenecut, B-adapt, keyframe rules,
 * cuTree chain walk) runs here on the host as scalar code; every pixel / block operation is a
 * batch of jobs sent to the GPU engine (libx265cu.so).  There is no CPU path for those jobs.
 *
 * Difference in mechanism, not in results: the reference computes motion searches and frame
 * costs on demand (or in thread-pool batches).  Here they are computed eagerly, in large GPU
 * batches, as soon as the frames involved are resident ("speculation"), and PUBLISHED into the
 * Lowres view in exactly the order the reference's control flow would have computed them.  An L0
 * search has two variants (run inside a P estimate or inside a B estimate, which differ by the
 * zero-MV skip rule, slicetype.cpp:4165-4181); both are kept on the device and the first touch
 * selects the one the reference would hold.
 */
#ifndef X265CU_LOOKAHEAD_H
#define X265CU_LOOKAHEAD_H

#include <stdint.h>
#include <stddef.h>
#include <deque>
#include <map>
#include <vector>
#include "x265cu.h"

namespace x265cu {

enum { TYPE_AUTO = 0, TYPE_IDR = 1, TYPE_I = 2, TYPE_P = 3, TYPE_BREF = 4, TYPE_B = 5 };   /* x265.h:572-577 */
enum { B_ADAPT_NONE = 0, B_ADAPT_FAST = 1, B_ADAPT_TRELLIS = 2 };                           /* x265.h:558-560 */
enum { BFRAME_MAX = 16, LOOKAHEAD_MAX = 250 };                                              /* x265.h:569,101 */
enum { LOWRES_COST_MASK = (1 << 14) - 1, LOWRES_COST_SHIFT = 14 };                          /* slicetype.h:41-42 */

inline bool isTypeI(int t) { return t == TYPE_I || t == TYPE_IDR; }
inline bool isTypeB(int t) { return t == TYPE_B || t == TYPE_BREF; }

/* The x265_param fields the lookahead reads (same names as source/x265.h). */
struct LookaheadParam
{
    int sourceWidth, sourceHeight, internalBitDepth, maxCUSize;
    uint32_t fpsNum, fpsDenom;
    int bframes, lookaheadDepth, bFrameAdaptive, bBPyramid, bFrameBias;
    int scenecutThreshold; double scenecutBias;   /* bias already divided by 100 (encoder.cpp:3948) */
    int keyframeMax, keyframeMin, bOpenGOP, bIntraRefresh;
    int radl;            /* --radl: leading pictures kept in front of a scene-cut IDR of a closed GOP */
    int gopLookahead;    /* --gop-lookahead: a keyframe due at the GOP boundary may wait this many frames for a scene cut */
    int bEnableWeightedPred, bEnableWeightedBiPred;
    int bEnableTemporalSubLayers;   /* --temporal-layers: 0 / 1, or 2 (B-refs placed recursively, their costs pre-computed by
                                       compCostBref, slicetype.cpp:1755-1799), or 3 / 4 / 5 (fixed random-access mini-GOPs of
                                       4 / 8 / 16 pictures, :2061-2325; Encoder::configure then sets bframes 3 / 7 / 15, b-adapt 0) */
    int bHistBasedSceneCut;         /* --hist-scenecut (8-bit only): scene cuts from per-segment histogram differences instead of
                                       the cost-based test (slicetype.cpp:3057-3216) */
    int bEnableHME;      /* --hme: hierarchical motion estimation, levels 0 (1/16 resolution) and 1 (lowres) of the lookahead's searches
                            (slicetype.cpp:4040-4048, 4083-4183); level 2 is the main encoder's */
    int hmeSearchMethod[2], hmeRange[2];    /* per level: X265_DIA/HEX/UMH/STAR/FULL_SEARCH (0..3, 5; sea is refused) and range */
    int bEnableFades;    /* --fades: mark the frame that ends a fade-in and code it as a keyframe (slicetype.cpp:1861-1906, 1972) */
    int lookaheadSlices;
    int maxNumReferences;
    struct
    {
        int aqMode; double aqStrength; int cuTree; double qCompress; int qgSize;
        int vbvBufferSize, vbvMaxBitrate, rateControlMode;
    } rc;
    /* Emulated ThreadPool::m_numWorkers.  The reference's results depend on it only through
     * m_bBatchMotionSearch / m_bBatchFrameCosts (slicetype.cpp:1024,1033,2693,2733), which decide
     * in which context an L0 search is first performed.  0 = "no pool" behaviour. */
    int poolWorkers;
    int device;          /* CUDA device ordinal */
    int extraSlots;      /* frames the caller may hold after getDecidedPicture before release */
    int speculate;       /* 1 (default): at every decision, one asynchronous GPU batch with the searches and costs of the
                            frames that arrived since the last one (those beyond the window included, see asyncDepth);
                            2: streaming -- one batch per frame, enqueued by addPicture (smaller launches: lower latency,
                            lower throughput); 0: on-demand jobs only.  Results are identical in all three */
    int shardCount;      /* > 1: this stream is sharded over that many ranks (x265cu_shard_config on the engine): frame poc is
                            computed by rank poc % shardCount, every rank takes the same decisions */
    int batchMin;        /* per-decision mode: frames beyond the window are batched once this many have gathered
                            (0 = asyncDepth / 2) */
    int pendingMax;      /* streaming + weightp: frames that may wait for their pixel sums before addPicture blocks on them */
    int asyncDepth;      /* extra frames of input delay before a decision is taken (0 = the reference's trigger).  The
                            decision analyses the same frames either way; the GPU gets that many frames of slack */
    int csvLogLevel;     /* x265_param::csvLogLevel */
    int numRowsPerSlice; /* > 0: the reference's m_numRowsPerSlice, taken as is (see la_capi.h) */
    int pinHost;         /* page-lock the picture buffers handed to addPicture the first time each is seen (they must then
                            stay allocated until destroy(); meant for an encoder's recycled PicYuv buffers) */
};
void lookaheadParamDefault(LookaheadParam* p);   /* x265_param_default + preset medium, param.cpp:164-349 */

class Lookahead;

/* Host view of one frame's lookahead state: the reference's Lowres (common/lowres.h:171-263),
 * minus pixel planes and big arrays, which stay on the device until fetched. */
struct Lowres
{
    int     frameNum, sliceType, sliceTypeReq, leadingBframes;
    bool    bScenecut, bKeyframe, bLastMiniGopBFrame, bIsFadeEnd;
    double  ipCostRatio;
    int64_t costEst[BFRAME_MAX + 2][BFRAME_MAX + 2];
    int64_t costEstAq[BFRAME_MAX + 2][BFRAME_MAX + 2];
    bool    rowSatdsValid[BFRAME_MAX + 2][BFRAME_MAX + 2];   /* rowSatds[i][j][0] != -1 */
    int     intraMbs[BFRAME_MAX + 2];
    int64_t satdCost;
    uint64_t wp_ssd[3], wp_sum[3];
    double  frameVariance;       /* --fades (slicetype.cpp:697-712) */
    /* --hist-scenecut: the picture statistics (collectPictureStatistics); hist points into Frame::m_hist */
    const x265cu_hist_stats* hist;
    bool    bHistScenecutAnalyzed;
    double  weightedCostDelta[BFRAME_MAX + 2];
    int     plannedType[LOOKAHEAD_MAX + 1];
    int64_t plannedSatd[LOOKAHEAD_MAX + 1];
    int     indB;
    int     rcD0, rcD1;      /* the (b - p0, p1 - b) estimate the last getEstimatedPictureCost read; -1 = none yet */
    int     rcPlanD0, rcPlanD1;  /* the estimate slicetypeDecide pre-computed for rate control (slicetype.cpp:2378-2427) and whose
                                    cuTree-adjusted cost was enqueued ahead of time; -1 = none */

    /* publication state: which device store holds what the reference would hold */
    int     mvStore[2][BFRAME_MAX + 2];                      /* -1 = lowresMvs[l][d][0].x == 0x7FFF */
    int     costStore[BFRAME_MAX + 2][BFRAME_MAX + 2];       /* -1 = never computed */
    /* weightp state per L0 distance: 0 unknown, 1 analysed/no weight, 2 weighted, 3 = searches and costs were enqueued
     * ASSUMING no weight because the pixel sums of the pair were not on the host yet (Lookahead::verifyWeights) */
    int     weightState[BFRAME_MAX + 2];
    int     wScale[BFRAME_MAX + 2], wDenom[BFRAME_MAX + 2], wOffset[BFRAME_MAX + 2];

    /* device-side bookkeeping.  A search "kind" is context + 3 * sliced: 0 = L0 in P context, 1 = L0 in B context,
     * 2 = L1, 3..5 = the same run as cooperative slices (only with LookaheadParam::lookaheadSlices + b-adapt 2 + pool,
     * where a search is sliced or not depending on whether a thread-pool batch touched it first).  A cost "variant"
     * names the searches it read: (L0 context) + 2 * (L0 sliced) + 4 * (L1 sliced). */
    int     slot;
    bool    statsFetched;
    uint8_t haveSearch[6][BFRAME_MAX + 2];                   /* kind x dist computed on (or queued for) the device */
    /* L0 distance d, per slicedness: 0 = unknown, 1 = the B-context search never applied the zero-MV skip, so the
     * P-context search is the same search and its store aliases the B-context one; 2 = the two variants differ */
    uint8_t l0Alias[2][BFRAME_MAX + 2];
    uint8_t flagFetched[2][BFRAME_MAX + 2];
    uint8_t haveCost[BFRAME_MAX + 2][BFRAME_MAX + 2][8];     /* cost variant computed on (or queued for) the device */
    uint8_t resultFetched[BFRAME_MAX + 2][BFRAME_MAX + 2][8];
    x265cu_cost_result result[BFRAME_MAX + 2][BFRAME_MAX + 2][8];
};

struct Frame
{
    int      m_poc;
    int64_t  m_pts, m_reorderedPts;
    bool     m_gopIdSet;      /* this decision wrote m_gopId (the reference leaves it alone for the B frames of a split mini-GOP, :2152) */
    int      m_gopOffset, m_gopId, m_tempLayer;     /* Frame::m_gopOffset / m_gopId / m_tempLayer (--temporal-layers 3..5: position in the
                                                       random-access structure, which structure, temporal layer; the DPB reads them) */
    Lowres   m_lowres;
    bool     m_lowresInit;
    bool     m_speculated;
    bool     m_released;      /* caller is done with it */
    bool     m_inUse;
    const void* m_planes[3];  /* caller's picture (valid until the frame's upload completed) */
    std::vector<x265cu_hist_stats> m_hist;   /* 0 or 1 entries (--hist-scenecut) */
    int      m_strideY, m_strideC;
    Lookahead* m_owner;
};

class Lookahead
{
public:
    explicit Lookahead(const LookaheadParam& param);
    ~Lookahead();

    /* same surface as the reference (slicetype.h:214-227) */
    bool    create();
    void    destroy();
    void    stopJobs() {}
    /* the reference takes a Frame the encoder filled; here the picture is handed over directly.
     * sliceType: Lookahead::addPicture's argument (the first-pass type of a 2-pass encode, else AUTO);
     * sliceTypeReq: x265_picture::sliceType as the application forces it (Frame::m_lowres.sliceTypeReq, encoder.cpp:1714).
     * Returns NULL (and sets the error string) when no frame slot is free (caller holds too many unreleased frames). */
    Frame*  addPicture(const void* y, const void* u, const void* v, int strideY, int strideC,
                       int64_t pts, int sliceType, int sliceTypeReq = TYPE_AUTO);
    void    flush();
    Frame*  getDecidedPicture();
    /* RateControl's entry point (slicetype.cpp:1327-1439).  The reference derives p0/p1 from the
     * slice's reference lists; the caller passes the POC distances instead (0 = none). */
    void    getEstimatedPictureCost(Frame* curFrame, Frame* ref0, Frame* ref1);
    void    getEstimatedPictureCost(Frame* curFrame, int d0, int d1);   /* POC distances to the references, 0 = none */
    /* its VBV half (slicetype.cpp:1387-1436), see lookahead.cpp; arrays of vbvRows() / geometry().ncu entries */
    bool    getVbvRowCosts(Frame* curFrame, int pirStartCol, int pirEndCol, uint32_t* satdForVbv, uint32_t* intraSatdForVbv,
                           uint16_t* lowresCostForRc, int32_t* intraCostScaled);
    int     vbvRows() const { return (m_param.sourceHeight + m_param.maxCUSize - 1) / m_param.maxCUSize; }
    int     findSliceType(int poc);
    void    releaseFrame(Frame* f);        /* encoder is done with the frame (DPB recycle) */

    /* host mirrors of device-resident Lowres arrays, in the reference's layout */
    bool    fetchMvs(Frame* f, int list, int dist, int32_t* mvXY, int32_t* mvCosts);  /* false + x=0x7FFF if unsearched */
    /* --hme: Lowres::lowerResMvs / lowerResMvCosts[list][dist] (the level-0 search behind a published lowres search) */
    bool fetchHmeMvs(Frame* f, int list, int dist, int32_t* mvXY, int32_t* mvCosts);
    bool    fetchCosts(Frame* f, int d0, int d1, uint16_t* lowresCosts, int32_t* rowSatds);
    bool    fetchFrame(Frame* f, const x265cu_frame_out* out);
    bool    mirror(Frame* f, const x265cu_mirror_request* req, int64_t* ticket);   /* asynchronous, x265cu_mirror_enqueue */

    /* sharded stream: this instance's rank.  Every rank takes the same slice-type decisions; only rank 0 is the decision rank
     * whose output is used, so only it runs cuTree and the rate-control costs (qp offsets never feed back into decisions) */
    void    setShardRank(int rank) { m_shardRank = rank; }
    const x265cu_geometry& geometry() const { return m_geom; }
    x265cu_ctx* engine() { return m_ctx; }
    const char* lastError() const { return m_error; }
    bool    ok() const { return !m_failed; }

    /* host-side wall-clock per phase (seconds), for bench.py: 0 pre-lookahead wait, 1 weightp, 2 enqueue,
     * 3 result wait, 4 decisions (host logic incl. cuTree enqueue), 5 whole slicetypeDecide, 6 calls,
     * 7 speculation inside addPicture (streaming mode; its weightp / enqueue shares are also in 1 / 2),
     * 8 getEstimatedPictureCost, 9 host mirrors (fetchMvs / fetchCosts / fetchFrame) */
    double  m_timers[10];

    LookaheadParam m_param;
    bool    m_filled;
    int     m_inputCount;

private:
    /* reference state (slicetype.h:155-203) */
    std::deque<Frame*> m_inputQueue, m_outputQueue;
    Lowres* m_lastNonB; Frame* m_lastNonBFrame;
    int     m_8x8Width, m_8x8Height, m_8x8Blocks, m_cuCount;
    int     m_lastKeyframe, m_fullQueueSize;
    /* --hist-scenecut state (slicetype.h:196-203) */
    uint32_t m_accHistDiffRunningAvg[4][4], m_accHistDiffRunningAvgCb[4][4], m_accHistDiffRunningAvgCr[4][4];
    bool    m_resetRunningAvg;
    uint32_t m_segmentCountThreshold;
    bool    histBasedScenecut(Lowres** frames, int p0, int p1, int numFrames);
    bool    detectHistBasedSceneChange(Lowres** frames, int p0, int p1, int p2);
    /* --fades state (slicetype.h:190-196) */
    double  m_frameVariance[BFRAME_MAX + 4];
    bool    m_isFadeIn;
    uint64_t m_fadeCount;
    int     m_fadeStart;
    bool    m_isSceneTransition, m_bBatchMotionSearch, m_bBatchFrameCosts, m_bAdaptiveQuant, m_extendGopBoundary;
    double  m_cuTreeStrength;
    int     m_rowsPerSlice;        /* cooperative search slices (slicetype.cpp:1047-1059); 0 = none */
    bool    m_dualSlicing;         /* sliced and unsliced variants of a search can both be needed (b-adapt 2 + pool) */
    bool    m_inBatch;             /* replaying one of the reference's thread-pool batches (:2668-2736) */
    int     m_costVariants;        /* 2, or 8 with dual slicing */

    /* engine */
    x265cu_ctx* m_ctx;
    x265cu_geometry m_geom;
    std::vector<uint16_t> m_mvcost;
    std::vector<Frame*> m_pool;           /* one per slot */
    std::vector<Frame*> m_resident;       /* frames with live slots, by arrival */
    std::deque<Frame*>  m_pendingSpec;    /* arrived, searches / costs not enqueued yet */
    std::vector<std::pair<Frame*, int> > m_unverified;   /* (frame, L0 distance) in weightState 3 */
    std::map<const void*, bool> m_pinned; /* caller buffers page-locked by pinHost (true = registered) */
    int     m_pocNext;
    int     m_shardRank;
    bool    m_shardDecouple;       /* sharded stream: cut batches without waiting for the pixel sums, like the unsharded path, but
                                      settle the assumed weights at fixed points only (every rank must cut identical batches) */
    bool    m_failed; char m_error[256];

    /* decision logic (same names as the reference) */
    void    slicetypeDecide();
    void    slicetypeAnalyse(Lowres** frames, Frame** fr, bool bKeyframe);
    bool    scenecut(Lowres** frames, int p0, int p1, bool bRealScenecut, int numFrames);
    bool    scenecutInternal(Lowres** frames, int p0, int p1, bool bRealScenecut);
    void    slicetypePath(Lowres** frames, int length, char (*best_paths)[LOOKAHEAD_MAX + 1]);
    int64_t slicetypePathCost(Lowres** frames, char* path, int64_t threshold);
    void    cuTree(Lowres** frames, int numframes, bool bIntra);
    void    estimateCUPropagate(Lowres** frames, double avgDuration, int p0, int p1, int b, int referenced);
    void    cuTreeFinish(Lowres* frame, double averageDuration, int ref0Distance);
    int64_t frameCostRecalculate(Lowres** frames, int p0, int p1, int b);
    void    vbvLookahead(Lowres** frames, int numFrames, int keyframe);
    int64_t vbvFrameCost(Lowres** frames, int p0, int p1, int b);
    void    placeBref(Frame** list, int start, int end, int num, int* brefs);
    void    compCostBref(Lowres** frames, int start, int end, int num);   /* :1780-1799 */
    void    decideTemporalLayers(Frame** list, Lowres** frames, Frame** fr, int bframes, int brefs, int& maxSearch);   /* :2061-2325 */

    /* CostEstimateGroup (slicetype.h:261-326) folded in */
    int64_t singleCost(Lowres** frames, int p0, int p1, int b, bool bIntraPenalty = false);
    int64_t estimateFrameCost(Lowres** frames, int p0, int p1, int b, bool bIntraPenalty);

    /* device orchestration */
    void    preLookahead(const std::vector<Frame*>& fr);
    void    speculateFrames(const std::vector<Frame*>& fresh, int onlyDist = 0);
    void    verifyWeights(int mustPoc);
    void    drainPending(size_t keep, int mustPoc);
    void    resolveAlias(const std::vector<Lowres*>& who);
    void    enqueueCosts(int l0kind, bool conditional);
    void    launchJobs();
    int     effKind(const Lowres* l, int d0, int kind) const { return (kind % 3 == 0 && l->l0Alias[kind / 3][d0] == 1) ? kind + 1 : kind; }
    /* cost variant of the searches (l0kind, l1kind) and the cost store it lives in */
    int     costVariant(int l0kind, int l1kind) const { return m_dualSlicing ? (l0kind % 3) + 2 * (l0kind / 3) + (l1kind >= 3 ? 4 : 0) : l0kind; }
    int     costStoreOf(int d0, int d1, int variant) const { return (d0 * m_geom.nb + d1) * m_costVariants + variant; }
    /* does a search first needed right now run as cooperative slices (slicetype.cpp:4004)? */
    int     sliceNow() const { return m_dualSlicing && !m_inBatch ? 1 : 0; }
    int     jobSliced(int kind) const { return m_rowsPerSlice && (!m_dualSlicing || kind >= 3) ? 1 : 0; }
    void    addSearch(Lowres* fenc, Lowres* ref, int kind, int d, int condStore = -1);
    void    addCost(Lowres* b, Lowres* p0, Lowres* p1, int d0, int d1, int l0kind, int l1kind, int condStore = -1);
    void    weightsAnalyseBatch(const std::vector<std::pair<Lowres*, Lowres*> >& pairs);
    void    ensureEstimate(Lowres* fenc, Lowres* ref0, Lowres* ref1, int d0, int d1, int l0kind, int l1kind);
    void    fetchResults(const std::vector<Lowres*>& who, int maxPoc);
    Frame*  frameOfPoc(int poc);
    void    recycle();
    void    initLowres(Frame* f, int poc);
    bool    check(int status, const char* what);
    void    fail(const char* what);

    std::vector<x265cu_search_job> m_searchJobs;
    std::vector<x265cu_cost_job>   m_costJobs;
};

void buildMvCostTable(std::vector<uint16_t>& table, int half, int depth);
int  lookaheadLambda(int depth);

} // namespace x265cu
#endif
