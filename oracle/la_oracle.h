/* la_oracle.h -- CPU restatement of the reference lookahead's block-level arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (x265-amod_b200/) may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg use it, and only as the checker.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against the
 * unmodified reference compiled from /root/reference (oracle/_ref, see Makefile.ref and
 * ref_harness.cpp) by tests/test_oracle_vs_ref.py, and against the committed golden
 * vectors under tests/golden/ (generated from that reference build by
 * tests/golden/make_golden.py) when the reference build is absent.
 *
 * Plain C99, scalar, written for clarity.  Compiled twice: -DOR_DEPTH=8 (pixel =
 * uint8_t) and -DOR_DEPTH=10 (pixel = uint16_t).  Citations are file:line under
 * /root/reference/source/. */
#ifndef LA_ORACLE_H
#define LA_ORACLE_H
#include <stdint.h>

#ifndef OR_DEPTH
#define OR_DEPTH 8
#endif
#if OR_DEPTH == 8
typedef uint8_t or_pixel;
#else
typedef uint16_t or_pixel;
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* Lowres geometry (common/lowres.cpp:72-97, common/picyuv.cpp:84-99) */
typedef struct
{
    int32_t picW, picH;     /* full-res luma size handed to the lookahead */
    int32_t w, h;           /* lowres plane size, rounded up to 8 */
    int32_t bw, bh, ncu;    /* 8x8 block grid */
    int32_t mx, my;         /* plane margins */
    int32_t stride;         /* plane stride in pixels */
    int32_t planeLines;     /* h + 2*my */
    int64_t planeSize;      /* stride * planeLines */
    int64_t padOffset;      /* my*stride + mx */
    int32_t rowsPerSlice;   /* Lookahead::m_numRowsPerSlice when the searches run as cooperative slices
                               (slicetype.cpp:1047-1059, 3957-3968); 0 = whole frame */
    int32_t qg8;            /* 1 = qg-size 8: AQ on 8x8 full-res blocks, 4 * ncu qp-offset entries (lowres.cpp:86-89) */
    /* --hme level 0: the 1/16-resolution planes (lowres.cpp:165-183, 378-388) and their block grid (slicetype.cpp:998-999) */
    int32_t w4, h4;         /* plane size: w / 2, h / 2 */
    int32_t bw4, bh4;       /* Lookahead::m_4x4Width / m_4x4Height */
    int32_t mx4, my4;       /* margins: mx / 2, my / 2 */
    int32_t stride4;        /* stride / 2 */
    int64_t planeSize4;     /* planeSize / 2 */
    int64_t padOffset4;     /* padOffset / 2 */
} or_geom;

int  or_depth(void);
void or_geom_init(or_geom* g, int picW, int picH, int maxCUSize);

/* lambda and mv-cost table of the lookahead (encoder/bitcost.cpp:30-54,98-113) */
int  or_lookahead_lambda(void);
void or_build_mvcost(uint16_t* table /* [2*half+1], centre at half */, int half);

int  or_exp2fix8(double x);                 /* common/common.cpp:96-103 */
int  or_sad8x8(const or_pixel* a, int sa, const or_pixel* b, int sb);   /* pixel.cpp:40-56 */
int  or_satd8x8(const or_pixel* a, int sa, const or_pixel* b, int sb);  /* pixel.cpp:239-297 */

/* Lowres::init pixel work: downscale + 4 hpel planes + border extension
 * (common/lowres.cpp:367-376, pixel.cpp:605-628,1044-1058).  `buf` is 4*planeSize pixels,
 * zero-initialised by the caller; the source is addressed with replicate clamping, which is
 * what PicYuv::copyFromPicture's padding (picyuv.cpp:261-285,480-512) amounts to. */
void or_lowres_init(const or_geom* g, const or_pixel* srcY, int srcStride, or_pixel* buf);

/* calcAdaptiveQuantFrame, aq-mode 0..5 (encoder/slicetype.cpp:452-713; modes 4 / 5 with edgeFilter / edgeDensityCu, :98-258); g->qg8 selects 8x8 blocks, in which case the
 * three output arrays hold 4 * ncu entries (+ a row of slack for the running index) and must come in zeroed.
 * Chroma may be NULL (treated as 4:0:0). */
void or_aq_frame(const or_geom* g, const or_pixel* y, int strideY, const or_pixel* u, const or_pixel* v,
                 int strideC, int aqMode, double aqStrength, int bWeightP,
                 double* qpAqOffset, double* qpCuTreeOffset, int32_t* invQscaleFactor,
                 uint32_t* blockEnergy /* ncu or NULL */, uint64_t wp_ssd[3], uint64_t wp_sum[3]);

/* --fades, the tail of calcAdaptiveQuantFrame (encoder/slicetype.cpp:697-712): returns Lowres::frameVariance and applies the
 * second acEnergyCu pass' side effect to wp_ssd / wp_sum; call after or_aq_frame */
double or_fade_variance(const or_geom* g, const or_pixel* y, int strideY, const or_pixel* u, const or_pixel* v, int strideC,
                        int bWeightP, uint64_t wp_ssd[3], uint64_t wp_sum[3]);

/* --hist-scenecut: the per-frame picture statistics of LookaheadTLD::collectPictureStatistics (encoder/slicetype.cpp:1441-1724)
 * over the quarter-sampled luma (common/lowres.cpp:35-51, 392-402) and the full-res chroma.  8-bit only: the reference
 * indexes 256 bins with the sample value.  Same layout as x265cu_hist_stats (include/x265cu.h). */
typedef struct
{
    uint32_t histogram[4][4][3][256];   /* Lowres::picHistogram[segment x][segment y][plane][bin] */
    uint8_t  avgIntensitySeg[4][4][3];  /* Lowres::averageIntensityPerSegment (the values are uint8_t casts) */
    uint8_t  avgIntensity[3];
    uint8_t  pad0;
    uint16_t picAvgVariance[3];         /* picAvgVariance, picAvgVarianceCb, picAvgVarianceCr */
    uint16_t pad1;
} or_hist_stats_t;
void or_hist_stats(const or_geom* g, const or_pixel* y, int strideY, const or_pixel* u, const or_pixel* v, int strideC,
                   const or_pixel* plane0 /* lowresPlane[0] */, or_hist_stats_t* out);

/* lowresIntraEstimate (encoder/slicetype.cpp:715-824) */
void or_intra_estimate(const or_geom* g, const or_pixel* plane0 /* lowresPlane[0] */,
                       const int32_t* invQscaleFactor /* may be NULL */,
                       int32_t* intraCost, uint8_t* intraMode, uint16_t* lowresCosts00,
                       int32_t* rowSatds00, int64_t* costEst00, int64_t* costEstAq00);

/* One list's motion search over the whole frame, reverse raster order, exactly the search
 * half of CostEstimateGroup::estimateCUCost (slicetype.cpp:4114-4183) + MotionEstimate::
 * motionEstimate (motion.cpp:764-1594, HEX + lowres subpel).  refPlanes[4] point at
 * lowresPlane[0..3] of the (possibly weighted) reference.  bBidir selects the B-frame
 * zero-MV skip rule (slicetype.cpp:4165-4181).  rowsPerSlice <= 0 means no slices. */
void or_search_list(const or_geom* g, const or_pixel* fencPlane0, const or_pixel* const refPlanes[4],
                    const uint16_t* mvcost /* centre */, int bBidir,
                    int32_t* mvs /* ncu*2 */, int32_t* mvCosts /* ncu */, int32_t* skipCount /* may be NULL */);

/* --hme (x265_param::bEnableHME): the four 1/16-resolution planes of a frame from its lowresPlane[0] (lowres.cpp:378-388) */
void or_lowerres_init(const or_geom* g, const or_pixel* plane0 /* lowresPlane[0], borders extended */, or_pixel* buf4 /* 4 * planeSize4, zeroed */);

/* The same search with --hme: `level` 0 searches the 1/16-resolution planes on the m_4x4 grid (fencPlane0 / refPlanes are then
 * lowerResPlane[0..3], never weighted, slicetype.cpp:4083-4094), level 1 is the ordinary lowres search with one more predictor,
 * twice the level-0 vector of the block's 16x16 parent when that search cost more than zero (:4142-4145; hmeMvs / hmeMvCosts =
 * the level-0 results of the same list and distance).  `method` is X265_DIA_SEARCH (0), X265_HEX_SEARCH (1) or X265_UMH_SEARCH (2)
 * (hmeSearchMethod[level], motion.cpp:842), `merange` hmeRange[level] (:4170).  Whole-frame searches only: with cooperative
 * slices the reference's two levels race (see la_oracle.c). */
void or_search_list_hme(const or_geom* g, int level, const or_pixel* fencPlane0, const or_pixel* const refPlanes[4],
                        const uint16_t* mvcost /* centre */, int bBidir, int method, int merange,
                        const int32_t* hmeMvs /* level 1: bw4*bh4*2 */, const int32_t* hmeMvCosts,
                        int32_t* mvs, int32_t* mvCosts, int32_t* skipCount /* may be NULL */);

/* Cost half of estimateCUCost + the sums of estimateFrameCost (slicetype.cpp:4187-4248,
 * 4050-4067).  For a P estimate pass ref1Planes = NULL. */
void or_frame_cost(const or_geom* g, const or_pixel* fencPlane0,
                   const or_pixel* const ref0Planes[4], const or_pixel* const ref1Planes[4],
                   const int32_t* mvs0, const int32_t* mvCosts0,
                   const int32_t* mvs1, const int32_t* mvCosts1,
                   const int32_t* intraCost, const int32_t* invQscaleFactor /* may be NULL */,
                   uint16_t* lowresCosts, int32_t* rowSatds,
                   int64_t* costEst, int64_t* costEstAq, int32_t* intraMbs);

/* weightCostLuma / weight_pp_c (slicetype.cpp:826-859, pixel.cpp:518-541) */
void or_weight_planes(const or_geom* g, const or_pixel* refBuf, or_pixel* dstBuf, int nPlanes,
                      int scale, int denom, int offset);
uint32_t or_weight_cost_luma(const or_geom* g, const or_pixel* fencPlane0, const or_pixel* refPlane0,
                             const int32_t* intraCost);
/* weightsAnalyse (slicetype.cpp:879-980): returns 1 and fills (scale,denom,offset,costDelta)
 * when a weight is accepted; wbuf is 4*planeSize scratch that receives the weighted planes. */
int or_weights_analyse(const or_geom* g, const or_pixel* fencPlane0, const or_pixel* refBuf,
                       const int32_t* intraCost, const uint64_t fenc_wp_ssd0, const uint64_t fenc_wp_sum0,
                       const uint64_t ref_wp_ssd0, const uint64_t ref_wp_sum0,
                       or_pixel* wbuf, int* scale, int* denom, int* offset, double* costDelta);

/* estimateCUPropagate for one (p0,p1,b) (slicetype.cpp:3502-3604 + pixel.cpp:931-957) */
void or_cutree_propagate(const or_geom* g, const int32_t* intraCost, const uint16_t* lowresCosts,
                         const int32_t* invQscaleFactor, const int32_t* mvs0, const int32_t* mvs1,
                         uint16_t* propagateB, uint16_t* refCost0, uint16_t* refCost1,
                         int referenced, int bipredWeight, double fpsFactor);
/* cuTreeFinish (slicetype.cpp:3750-3798, qg-size > 8 branch) */
void or_cutree_finish(const or_geom* g, const int32_t* intraCost, const int32_t* invQscaleFactor,
                      const uint16_t* propagateCost, const double* qpAqOffset, double* qpCuTreeOffset,
                      int fpsFactorFix8, double weightdelta, double cuTreeStrength);
/* frameCostRecalculate (slicetype.cpp:3802-3879, non-hevc-aq branch) */
int64_t or_frame_cost_recalc(const or_geom* g, const uint16_t* lowresCosts, const double* qpOffset,
                             int32_t* rowSatds);
/* qg-size 8: invQscaleFactor8x8 (slicetype.cpp:656-670) */
void or_invq8x8(const or_geom* g, const int32_t* invQ, int32_t* invQ8);
/* the VBV row aggregation of getEstimatedPictureCost (slicetype.cpp:1387-1436); qpOffset NULL = no scaling,
 * pirStart < 0 = no intra-refresh term */
void or_vbv_rows(const or_geom* g, const uint16_t* lowresCosts, const int32_t* intraCost, const double* qpOffset,
                 int scale, int pirStart, int pirEnd, int nRows, uint32_t* satdForVbv, uint32_t* intraSatdForVbv,
                 uint16_t* lowresCostForRc, int32_t* intraCostScaled);

#ifdef __cplusplus
}
#endif
#endif
