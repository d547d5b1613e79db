/* la_capi.h -- plain C surface of the host Lookahead (x265-amod_b200/host/lookahead.h) so that
 * bench.py, the tests and non-C++ callers can drive it with host buffers.  Thin: every call maps
 * 1:1 to a method of x265cu::Lookahead, which mirrors the reference's class Lookahead
 * (source/encoder/slicetype.h:151-259). */
#ifndef X265LA_CAPI_H
#define X265LA_CAPI_H
#include <stdint.h>
#include "x265cu.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    int32_t sourceWidth, sourceHeight, internalBitDepth, maxCUSize;
    int32_t fpsNum, fpsDenom;
    int32_t bframes, lookaheadDepth, bFrameAdaptive, bBPyramid, bFrameBias;
    int32_t scenecutThreshold;
    double  scenecutBias;            /* percent, like --scenecut-bias (divided by 100 internally) */
    int32_t keyframeMax, keyframeMin, bOpenGOP, bIntraRefresh;
    int32_t bEnableWeightedPred, bEnableWeightedBiPred;
    int32_t lookaheadSlices, maxNumReferences;
    int32_t aqMode; double aqStrength; int32_t cuTree; double qCompress; int32_t qgSize;
    int32_t vbvBufferSize, vbvMaxBitrate, rateControlMode;
    int32_t poolWorkers, device, extraSlots, speculate, pinHost;
    int32_t asyncDepth;              /* LookaheadParam::asyncDepth */
    int32_t pendingMax;              /* LookaheadParam::pendingMax; 0 = default */
    int32_t shardCount;              /* LookaheadParam::shardCount; 0 / 1 = not sharded */
    int32_t batchMin;                /* LookaheadParam::batchMin; 0 = asyncDepth / 2 */
    int32_t gopLookahead;            /* x265_param::gopLookahead */
    int32_t radl;                    /* x265_param::radl */
    int32_t csvLogLevel;             /* x265_param::csvLogLevel: >= 2 makes scenecut() record Lowres::ipCostRatio (slicetype.cpp:2998-3003) */
    int32_t numRowsPerSlice;         /* > 0: Lookahead::m_numRowsPerSlice as the reference's constructor derived it (it rewrites
                                        x265_param::lookaheadSlices afterwards, so deriving it twice is not idempotent); 0 = derive
                                        from lookaheadSlices */
    int32_t bEnableFades;            /* x265_param::bEnableFades (--fades) */
    int32_t bEnableTemporalSubLayers;/* x265_param::bEnableTemporalSubLayers (--temporal-layers): 0-2 */
    int32_t bHistBasedSceneCut;      /* x265_param::bHistBasedSceneCut (--hist-scenecut), 8-bit only */
    int32_t bEnableHME;              /* x265_param::bEnableHME (--hme); like the encoder, ignored below 540 lines (encoder.cpp:4400-4407) */
    int32_t hmeSearchMethod[2];      /* x265_param::hmeSearchMethod[0..1]: dia (0), hex (1), umh (2), star (3) or full (5) */
    int32_t hmeRange[2];             /* x265_param::hmeRange[0..1] */
} x265la_param;

typedef struct
{
    int32_t poc, sliceType, bScenecut, bKeyframe, bLastMiniGopBFrame, leadingBframes;
    int64_t pts, reorderedPts, satdCost;
    void*   handle;                  /* pass to the x265la_frame_* calls and x265la_release */
    /* --temporal-layers 3..5: Frame::m_gopOffset / m_gopId / m_tempLayer as slicetypeDecide left them (slicetype.cpp:2133-2320);
     * gopIdWritten == 0: the reference did not touch m_gopId of this frame (the caller's Frame keeps what it held) */
    int32_t gopOffset, gopId, tempLayer, gopIdWritten;
} x265la_frame_info;

void  x265la_param_default(x265la_param* p);
void* x265la_open(const x265la_param* p, char* err, int32_t errLen);   /* NULL on failure */
void  x265la_close(void* la);
int   x265la_get_geometry(void* la, x265cu_geometry* g);
const char* x265la_last_error(void* la);
/* Lookahead::addPicture; returns the frame handle or NULL.  sliceType: the argument of the reference's addPicture (first-pass
 * type of a 2-pass encode, else X265_TYPE_AUTO); sliceTypeReq: x265_picture::sliceType as an application forces it
 * (Encoder::encode stores it in Lowres::sliceTypeReq, encoder.cpp:1714) */
void* x265la_add_picture(void* la, const void* y, const void* u, const void* v, int32_t strideY, int32_t strideC,
                         int64_t pts, int32_t sliceType, int32_t sliceTypeReq);
void  x265la_flush(void* la);
/* Lookahead::getDecidedPicture; 1 = frame returned, 0 = none yet, <0 = error */
int   x265la_get_decided(void* la, x265la_frame_info* out);
/* Lookahead::getEstimatedPictureCost with explicit references (handles, may be NULL) */
int64_t x265la_estimated_picture_cost(void* la, void* frame, void* ref0, void* ref1);
/* the same with the POC distances to the list-0 / list-1 reference (0 = none) instead of their handles, for a caller
 * that has already released the reference frames */
int64_t x265la_estimated_picture_cost_dist(void* la, void* frame, int32_t d0, int32_t d1);
int     x265la_find_slice_type(void* la, int32_t poc);      /* Lookahead::findSliceType */
double  x265la_frame_ip_cost_ratio(void* la, void* frame);  /* Lowres::ipCostRatio */
/* the VBV half of getEstimatedPictureCost (slicetype.cpp:1387-1436) for the estimate the last
 * x265la_estimated_picture_cost of this frame read: x265la_vbv_rows() row sums, ncu scaled costs (NULL = skip); 0 = ok */
int   x265la_vbv_rows(void* la);
int   x265la_vbv_row_costs(void* la, void* frame, int32_t pirStartCol, int32_t pirEndCol, uint32_t* satdForVbv,
                           uint32_t* intraSatdForVbv, uint16_t* lowresCostForRc, int32_t* intraCostScaled);
/* what RateControl reads of the VBV lookahead (ratecontrol.cpp:2512-2577): Lowres::plannedSatd / plannedType (n entries
 * each, n <= 251) and indB */
int   x265la_frame_planned(void* la, void* frame, int64_t* plannedSatd, int32_t* plannedType, int32_t n, int32_t* indB);
void  x265la_release(void* la, void* frame);

/* published Lowres state of one frame, reference layout (common/lowres.h:171-263) */
int   x265la_frame_scalars(void* la, void* frame, int64_t* costEst /* nb*nb */, int64_t* costEstAq /* nb*nb */,
                           int32_t* intraMbs /* nb */, int32_t* rowSatdsValid /* nb*nb */,
                           uint64_t* wp_ssd /* 3 */, uint64_t* wp_sum /* 3 */, double* weightedCostDelta /* nb */);
int   x265la_frame_mvs(void* la, void* frame, int32_t list, int32_t dist, int32_t* mvXY, int32_t* mvCosts);  /* 1 ok, 0 unsearched */
/* --hme: the level-0 vectors / costs (Lowres::lowerResMvs / lowerResMvCosts) of a published search; 1 ok, 0 none */
int   x265la_frame_hme_mvs(void* la, void* frame, int32_t list, int32_t dist, int32_t* mvXY, int32_t* mvCosts);
int   x265la_frame_costs(void* la, void* frame, int32_t d0, int32_t d1, uint16_t* lowresCosts, int32_t* rowSatds);
int   x265la_frame_fetch(void* la, void* frame, const x265cu_frame_out* out);
/* The host mirror of a decided frame in one asynchronous request (x265cu_mirror_enqueue): every destination is a host array in
 * the reference's layout, NULL = skip.  lowresMvs[list][dist]: (x, y) int32 pairs = the reference's MV struct; lists the
 * lookahead never published are skipped and their bit stays clear in published[list] (the caller writes the 0x7FFF sentinel).
 * (d0, d1) names the estimate whose lowresCosts / rowSatds are wanted, d0 < 0 = none.  Returns 0 and a ticket for
 * x265la_mirror_wait.  Page-lock the destinations (x265la_pin) or the enqueue itself waits for the copies. */
typedef struct
{
    int32_t*  intraCost; double* qpAqOffset; double* qpCuTreeOffset; int32_t* invQscaleFactor; void* planes;
    int32_t*  lowresMvs[2][18];
    int32_t   d0, d1; uint16_t* lowresCosts; int32_t* rowSatds;
} x265la_mirror;
int   x265la_frame_mirror_async(void* la, void* frame, const x265la_mirror* m, uint32_t published[2], int64_t* ticket);
int   x265la_mirror_wait(void* la, int64_t ticket);
int   x265la_pin(void* la, void* ptr, uint64_t bytes);      /* cudaHostRegister through the engine; 0 = ok */
int   x265la_unpin(void* la, void* ptr);
/* --fades: Lowres::bIsFadeEnd (rate control resets on it, ratecontrol.cpp:1416) and Lowres::frameVariance */
int   x265la_frame_fade(void* la, void* frame, int32_t* bIsFadeEnd, double* frameVariance);
/* --hist-scenecut: Lowres::picAvgVariance{,Cb,Cr}, averageIntensity[3] and a checksum over picHistogram + averageIntensityPerSegment
 * (tests compare it with the same checksum of the reference's arrays); returns 0 when the frame has such statistics */
int   x265la_frame_hist(void* la, void* frame, int32_t* variance /* 3 */, int32_t* intensity /* 3 */, uint64_t* checksum);
/* weightp analysis outcome per L0 distance: state 0 = not analysed, 1 = no weight, 2 = weighted */
int   x265la_frame_weights(void* la, void* frame, int32_t* state, int32_t* scale, int32_t* denom, int32_t* offset /* nb each */);
/* host wall-clock per phase, seconds (see Lookahead::m_timers); reset != 0 clears them */
int   x265la_get_timers(void* la, double* t /* 10 */, int32_t reset);
/* the BitCost row of the lookahead QP as the host library builds it (bitcost.cpp:46-54, 98-113): entries [-half, +half], centre
 * at table[half].  For tests, and for a caller that wants to compare it with the encoder's own */
void  x265la_mvcost_table(int32_t depth, int32_t half, uint16_t* table);
x265cu_ctx* x265la_engine(void* la);
/* sharded stream: x265cu_shard_config on this Lookahead's engine (open it with x265la_param::shardCount = nranks) */
int   x265la_shard_config(void* la, int32_t rank, int32_t nranks, x265cu_exchange_fn exchange, void* user);

#ifdef __cplusplus
}
#endif
#endif
