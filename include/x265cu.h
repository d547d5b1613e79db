/* x265cu.h -- C ABI of the B200 lookahead engine (libx265cu.so).
 *
 * This is the drop-in boundary for the reference's lookahead hot path.  The reference has no
 * FFI for this path (the only operator-style interface is the per-8x8-block EncoderPrimitives
 * table, source/common/primitives.h:239-436); these entry points are what a CMake ENABLE_CUDA
 * build of the reference binds underneath the bodies listed per function below.  Host code
 * (x265-amod_b200/host/, or the patched reference) owns every decision;
 * the engine owns every pixel/block operation and keeps a whole lookahead window resident in HBM.
 *
 * Conventions: extern "C", opaque context, plain pointers and sizes, every call returns an
 * int status (0 = X265CU_OK, negative = error; x265cu_strerror() names it).  Nothing throws
 * (the reference is built -fno-exceptions, source/CMakeLists.txt:358-362).  A context is
 * bound to one Lookahead instance and one GPU; contexts are independent (abrEncApp runs
 * several Lookaheads per process, source/abrEncApp.cpp:510).  One caller thread at a time per
 * context.  There is NO CPU fallback: x265cu_create fails with X265CU_ERR_NO_DEVICE when no
 * sm_100 device is usable.
 *
 * Work is expressed as explicit batches of jobs.  Calls that enqueue work are asynchronous;
 * the calls documented as "synchronises" wait for the results they return.
 */
#ifndef X265CU_H
#define X265CU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X265CU_OK               0
#define X265CU_ERR_NO_DEVICE   -1   /* no usable CUDA device / driver */
#define X265CU_ERR_BAD_ARG     -2
#define X265CU_ERR_NO_MEMORY   -3
#define X265CU_ERR_CUDA        -4   /* a CUDA call or kernel failed; see x265cu_last_error */
#define X265CU_ERR_UNSUPPORTED -5   /* configuration outside the hot path (sea search under --hme, --fades with aq-mode 4 / 5, ...) */

typedef struct x265cu_ctx x265cu_ctx;

/* Geometry and constants fixed for the life of a Lookahead
 * (Lookahead::Lookahead, encoder/slicetype.cpp:982-1059; Lowres::create, common/lowres.cpp:72-251). */
typedef struct
{
    int32_t width, height;      /* x265_param::sourceWidth/Height after Encoder::configure padding */
    int32_t depth;              /* X265_DEPTH: 8, 10 or 12 (planes are uint8_t for 8, uint16_t otherwise) */
    int32_t max_cu_size;        /* x265_param::maxCUSize (plane margins, picyuv.cpp:87-88) */
    int32_t bframes;            /* x265_param::bframes; per-frame arrays are (bframes+2) wide */
    int32_t max_slots;          /* frame slots resident in HBM */
    int32_t qg_size;            /* x265_param::rc.qgSize: 8, 16, 32 or 64.  8 = AQ on 8x8 full-res blocks: the qp-offset arrays
                                   hold ncu_full = 4 * ncu entries, addressed like the reference (lowres.h:98-106) */
    int32_t aq_mode;            /* x265_param::rc.aqMode 0..5 (4 / 5 = edge: Gaussian + gradient edge map of the luma, slicetype.cpp:98-258;
                                   not together with fade_stats) */
    double  aq_strength;        /* x265_param::rc.aqStrength */
    int32_t need_aq;            /* Lookahead::m_bAdaptiveQuant (slicetype.cpp:1013-1017) */
    int32_t need_wp_stats;      /* bEnableWeightedPred || bEnableWeightedBiPred */
    int32_t lambda;             /* (int)x265_lambda_tab[X265_LOOKAHEAD_QP] */
    const uint16_t* mvcost;     /* BitCost row for X265_LOOKAHEAD_QP (bitcost.cpp:46-54): entries */
    int32_t mvcost_half;        /*   [-mvcost_half, +mvcost_half], centre at mvcost[mvcost_half]   */
    int32_t device;             /* CUDA device ordinal */
    int32_t rows_per_slice;     /* Lookahead::m_numRowsPerSlice, the height of a cooperative search slice
                                   (slicetype.cpp:1047-1059, 3957-3968: a slice's bottom row takes no predictors from
                                   the slice below); 0 = no slices.  Applied to the search jobs that ask for it */
    int32_t mv_store_kinds;     /* MV stores per slot = this * (bframes+2); 0 = 3 (L0 in P context, L0 in B context, L1).
                                   6 when sliced and unsliced variants of a search can both be needed */
    int32_t cost_variants;      /* cost stores per slot = this * (bframes+2)^2; 0 = 2 */
    int32_t fade_stats;         /* x265_param::bEnableFades: Lowres::frameVariance per frame and the second acEnergyCu pass'
                                   side effect on wp_ssd / wp_sum (slicetype.cpp:697-712).  Needs need_aq, qg-size > 8 */
    int32_t hist_stats;         /* x265_param::bHistBasedSceneCut: the per-frame picture statistics of collectPictureStatistics
                                   (slicetype.cpp:1441-1724), read back with x265cu_frame_hist_get.  8-bit only */
    int32_t hme;                /* x265_param::bEnableHME (--hme): every search job first searches the 1/16-resolution planes
                                   (lowres.cpp:378-388) on the m_4x4 block grid and feeds twice that vector to the lowres search as one
                                   more predictor (slicetype.cpp:4040-4048, 4142-4145).  The level-0 results stay on the device */
    int32_t hme_search[2];      /* x265_param::hmeSearchMethod[0..1] (level 2 is the main encoder's): X265_DIA_SEARCH (0),
                                   X265_HEX_SEARCH (1), X265_UMH_SEARCH (2), X265_STAR_SEARCH (3) or X265_FULL_SEARCH (5); X265_SEA (4) is refused */
    int32_t hme_range[2];       /* x265_param::hmeRange[0..1] */
    int32_t reserved[2];
} x265cu_config;

/* derived geometry, as Lowres::create computes it */
typedef struct
{
    int32_t low_width, low_height;   /* lowres plane size (multiples of 8) */
    int32_t bw, bh, ncu;             /* 8x8 block grid: m_8x8Width/Height/m_cuCount */
    int32_t stride, plane_lines;     /* plane stride (pixels) and lines incl. margins */
    int32_t margin_x, margin_y;
    int32_t nb;                      /* bframes + 2 */
    int32_t n_mv_stores;             /* MV stores per slot  = 3*nb (see x265cu_search_job::store) */
    int32_t n_cost_stores;           /* cost stores per slot = 2*nb*nb */
    int32_t ncu_full;                /* entries of qpAqOffset / qpCuTreeOffset / invQscaleFactor: ncu, or 4 * ncu with qg-size 8 */
} x265cu_geometry;

int  x265cu_device_count(void);
int  x265cu_create(const x265cu_config* cfg, x265cu_ctx** out);
void x265cu_destroy(x265cu_ctx* ctx);
int  x265cu_get_geometry(const x265cu_ctx* ctx, x265cu_geometry* out);
/* SMs of the two partitions the engine runs on (CUDA green contexts): `small_sms` for the short latency-critical kernels the
 * host waits for (cuTree, cost recalculation, weightp scores, mirrors), `large_sms` for the search / cost batches; 0 / 0 when
 * the device is not partitioned (driver without green contexts, or X265CU_GREEN=0 in the environment) */
int  x265cu_sm_partition(const x265cu_ctx* ctx, int32_t* small_sms, int32_t* large_sms);
const char* x265cu_strerror(int status);
const char* x265cu_last_error(const x265cu_ctx* ctx);

/* Page-lock / unlock a host picture buffer so uploads run at full PCIe rate (optional). */
int  x265cu_pin_host(x265cu_ctx* ctx, void* ptr, uint64_t bytes);
int  x265cu_unpin_host(x265cu_ctx* ctx, void* ptr);

/* ---- pre-lookahead: replaces the body of PreLookaheadGroup::processTasks
 * (slicetype.cpp:1726-1752): Lowres::init pixel work (lowres.cpp:367-376), calcAdaptiveQuantFrame
 * (slicetype.cpp:452-713) and lowresIntraEstimate (:715-824) for one frame.
 * Planes are `depth`-bit samples in uint8_t (depth 8) or uint16_t; strides in samples; the
 * picture is width x height (chroma 4:2:0, may be NULL for 4:0:0).  Asynchronous: the host
 * buffers must stay valid until the next synchronising call on this context. */
int  x265cu_frame_upload(x265cu_ctx* ctx, int32_t slot, const void* y, const void* u, const void* v,
                         int32_t stride_y, int32_t stride_c);

typedef struct
{
    int64_t  cost_est;      /* Lowres::costEst[0][0]   */
    int64_t  cost_est_aq;   /* Lowres::costEstAq[0][0] */
    uint64_t wp_ssd[3];     /* Lowres::wp_ssd (finalised, slicetype.cpp:681-694) */
    uint64_t wp_sum[3];     /* Lowres::wp_sum */
    double   frame_variance;/* Lowres::frameVariance (x265cu_config::fade_stats), else 0 */
} x265cu_frame_stats;
/* --hist-scenecut: Lowres::picHistogram / averageIntensityPerSegment / averageIntensity / picAvgVariance{,Cb,Cr} of one frame
 * (x265cu_config::hist_stats).  Synchronises with the frame's pre-lookahead only. */
typedef struct
{
    uint32_t histogram[4][4][3][256];       /* [segment x][segment y][plane][bin] */
    uint8_t  avg_intensity_seg[4][4][3];
    uint8_t  avg_intensity[3];
    uint8_t  pad0;
    uint16_t pic_avg_variance[3];
    uint16_t pad1;
} x265cu_hist_stats;
int  x265cu_frame_hist_get(x265cu_ctx* ctx, int32_t slot, x265cu_hist_stats* out);

/* 1 when the pre-lookahead of the frame in `slot` has finished (x265cu_frame_stats_get would not wait), 0 when it is
 * still running, negative on error.  Never blocks. */
int  x265cu_frame_ready(x265cu_ctx* ctx, int32_t slot);
/* synchronises (with the pre-lookahead of these frames only) */
int  x265cu_frame_stats_get(x265cu_ctx* ctx, const int32_t* slots, int32_t n, x265cu_frame_stats* out);

/* ---- asynchronous batches.  Search / cost jobs enqueued between batch_begin and batch_end run on one of the
 * engine's worker streams, concurrently with the batches opened before and after (a lowres search is a ~500-step
 * wavefront, so one frame's jobs cannot fill the GPU; a window of frames in flight can).  The engine orders a
 * batch after the pre-lookahead of every frame uploaded before batch_begin and after the searches whose MV stores
 * its cost jobs read; every call that reads a store waits for the batch that writes it, and a frame upload waits
 * for the batches that still use the slot's previous tenant.  Jobs enqueued outside begin/end form a batch of their
 * own.  Nothing here blocks the caller. */
int  x265cu_batch_begin(x265cu_ctx* ctx, int64_t* batch_id /* may be NULL */);
int  x265cu_batch_end(x265cu_ctx* ctx);
/* batches still open or running, counted up to 2 (never blocks; 0, 1, 2, or a negative status): lets the caller launch a smaller
 * batch early when the GPU would otherwise run dry waiting for a full one (it should hold the batch it works on and one behind it) */
int  x265cu_batches_in_flight(x265cu_ctx* ctx);

/* ---- one stream sharded over several GPUs (SURVEY 8e level 2): every (frame, list, distance) search and every
 * (p0, p1, b) estimate depends only on pixels, so the ranks of a job split them by source frame.  Every rank runs the
 * same host logic on the same pictures and holds every frame; a rank only COMPUTES the search / cost jobs of the
 * frames it owns (x265cu_slot_owner), and at the end of each batch the MV / cost stores written by their owners are
 * exchanged so that every rank holds all of them (what cuTree and the decisions read afterwards is then identical
 * everywhere; rank 0's output is the product).  The engine packs what it owns, calls `exchange` once per batch and
 * unpacks the rest; the callback provides the collective (NCCL broadcasts through torch.distributed in bench.py, gloo
 * in the CPU tests): for every root r with bytes[r] > 0 it must make bufs[r] on every rank equal rank r's bufs[r],
 * ordered on `cuda_stream` (a cudaStream_t).  All ranks call it with the same bytes[].  Returns 0 on success. */
typedef int (*x265cu_exchange_fn)(void* user, void* const* bufs, const uint64_t* bytes, int32_t nranks, void* cuda_stream);
int  x265cu_shard_config(x265cu_ctx* ctx, int32_t rank, int32_t nranks /* <= 8 */, x265cu_exchange_fn exchange, void* user);
int  x265cu_slot_owner(x265cu_ctx* ctx, int32_t slot, int32_t owner_rank);

/* ---- motion search: the search half of CostEstimateGroup::estimateCUCost (slicetype.cpp:
 * 4103-4183) + MotionEstimate::motionEstimate (motion.cpp:764-1594, HEX + lowres subpel) for one
 * (frame, list, distance) over the whole frame, reverse-raster dependency order preserved. */
typedef struct
{
    int32_t fenc_slot, ref_slot;
    int32_t bidir_ctx;      /* 1 when the reference would run this search inside a B estimate
                               (b < p1): enables the zero-MV skip rule, slicetype.cpp:4165-4181 */
    int32_t store;          /* MV store of fenc_slot that receives MVs + MV costs:
                               kind*nb + dist, kind 0 = L0 in P context, 1 = L0 in B context, 2 = L1
                               (3..5 = the same, sliced, when mv_store_kinds is 6) */
    int32_t weighted;       /* search the weighted copy of the reference (slicetype.cpp:4083,4128) */
    int32_t w_scale, w_denom, w_offset; /* WeightParam inputWeight/log2WeightDenom/inputOffset */
    int32_t cond_store;     /* -1 = always run.  Else the job runs only if the search held in MV store `cond_store`
                               of fenc_slot applied the zero-MV skip rule to at least one block; decided on the
                               device when the job starts, so the host need not wait for that search (a B-context
                               L0 search that never skipped IS the P-context search, see x265cu_search_flags_get) */
    int32_t sliced;         /* search as cooperative slices of x265cu_config::rows_per_slice rows (the reference runs a
                               search that way when it is first needed outside a thread-pool batch, slicetype.cpp:4004) */
} x265cu_search_job;
int  x265cu_search_batch(x265cu_ctx* ctx, const x265cu_search_job* jobs, int32_t n);
/* synchronises; flags[i] != 0 when the search stored in (slots[i], stores[i]) applied the zero-MV skip rule to at
 * least one block.  A B-context L0 search that never did is identical to the P-context search of the same
 * (frame, distance), which lets the host skip the second variant. */
int  x265cu_search_flags_get(x265cu_ctx* ctx, const int32_t* slots, const int32_t* stores, int32_t n, int32_t* flags);

/* ---- frame cost: the cost half of estimateCUCost (slicetype.cpp:4187-4248) and the sums of
 * estimateFrameCost (:4050-4062) for one (p0,p1,b).  P estimate: p1_slot == b_slot, l1_store < 0. */
typedef struct
{
    int32_t b_slot, p0_slot, p1_slot;
    int32_t l0_store, l1_store;   /* MV stores of b_slot to read */
    int32_t out;                  /* cost store of b_slot: (d0*nb + d1)*2 + variant */
    int32_t cond_store;           /* -1 = always; else like x265cu_search_job::cond_store (MV store of b_slot) */
} x265cu_cost_job;
int  x265cu_cost_batch(x265cu_ctx* ctx, const x265cu_cost_job* jobs, int32_t n);

typedef struct
{
    int64_t cost_est;       /* sum over interior blocks, before the B-frame scaling of :4064-4065 */
    int64_t cost_est_aq;
    int32_t intra_mbs;      /* P estimates only */
    int32_t reserved;
} x265cu_cost_result;
/* synchronises; result i is for (slots[i], outs[i]) */
int  x265cu_cost_results_get(x265cu_ctx* ctx, const int32_t* slots, const int32_t* outs, int32_t n,
                             x265cu_cost_result* res);

/* ---- weightp: LookaheadTLD::weightCostLuma (slicetype.cpp:826-859).  weighted == 0 scores the
 * plain reference; else weight_pp (pixel.cpp:518-541) is applied first.  synchronises. */
typedef struct
{
    int32_t fenc_slot, ref_slot;
    int32_t weighted, w_scale, w_denom, w_offset;
} x265cu_wcost_job;
int  x265cu_weight_cost_batch(x265cu_ctx* ctx, const x265cu_wcost_job* jobs, int32_t n, uint32_t* costs);

/* ---- cuTree: Lookahead::estimateCUPropagate (slicetype.cpp:3502-3604) with
 * primitives.propagateCost (pixel.cpp:931-957), cuTreeFinish (:3750-3798) and
 * frameCostRecalculate (:3802-3879). */
int  x265cu_cutree_reset(x265cu_ctx* ctx, int32_t slot);   /* memset(propagateCost, 0) */
int  x265cu_cutree_propagate(x265cu_ctx* ctx, int32_t b_slot, int32_t p0_slot, int32_t p1_slot,
                             int32_t cost_store, int32_t l0_store, int32_t l1_store,
                             int32_t referenced, int32_t bipred_weight, double fps_factor);
int  x265cu_cutree_finish(x265cu_ctx* ctx, int32_t slot, int32_t fps_factor_fix8, double weightdelta,
                          double cutree_strength);
/* synchronises; row_satds may be NULL */
int  x265cu_cost_recalc(x265cu_ctx* ctx, int32_t slot, int32_t cost_store, int32_t use_cutree_offsets,
                        int64_t* score, int32_t* row_satds);

/* The VBV half of Lookahead::getEstimatedPictureCost (slicetype.cpp:1387-1436): lowresCostForRc = the block costs of cost
 * store `cost_store` (0 = the intra estimate) and intraCost, both scaled by the block's qp offset (qp_source 0 = none,
 * 1 = qpAqOffset, 2 = qpCuTreeOffset), and their sums per CTU row of `ctu_rows_lowres` lowres block rows
 * (maxCUSize / 16): what the reference adds to FrameData::m_rowStat[].satdForVbv / intraSatdForVbv.  pir_start / pir_end:
 * the intra-refresh columns of a P slice (:1425-1427), -1 = none.  Outputs are host arrays (NULL = skip): n_rows sums each,
 * ncu scaled costs each.  The device copies of intraCost / lowresCosts are NOT modified (the reference rewrites its host
 * arrays in place; the caller's mirrors receive the same values).  synchronises */
int  x265cu_vbv_row_costs(x265cu_ctx* ctx, int32_t slot, int32_t cost_store, int32_t qp_source, int32_t ctu_rows_lowres,
                          int32_t pir_start, int32_t pir_end, int32_t n_rows, uint32_t* satd_for_vbv, uint32_t* intra_satd_for_vbv,
                          uint16_t* lowres_cost_for_rc, int32_t* intra_cost_scaled);

/* ---- host mirrors of Lowres fields (all synchronise; NULL pointers are skipped) */
typedef struct
{
    int32_t*  intra_cost;        /* ncu */
    uint8_t*  intra_mode;        /* ncu */
    double*   qp_aq_offset;      /* ncu_full */
    double*   qp_cutree_offset;  /* ncu_full */
    int32_t*  inv_qscale_factor; /* ncu_full */
    uint16_t* propagate_cost;    /* ncu */
    void*     planes;            /* 4 * stride * plane_lines samples: Lowres::buffer[0] */
    uint16_t* lowres_costs00;    /* ncu: lowresCosts[0][0] */
    int32_t*  row_satds00;       /* bh:  rowSatds[0][0] */
} x265cu_frame_out;
int  x265cu_fetch_frame(x265cu_ctx* ctx, int32_t slot, const x265cu_frame_out* out);
int  x265cu_fetch_mvs(x265cu_ctx* ctx, int32_t slot, int32_t store, int32_t* mv_xy /* ncu*2 */,
                      int32_t* mv_costs /* ncu */);
int  x265cu_fetch_costs(x265cu_ctx* ctx, int32_t slot, int32_t cost_store, uint16_t* lowres_costs /* ncu */,
                        int32_t* row_satds /* bh */);
/* --hme: the level-0 results behind MV store `store` (Lowres::lowerResMvs / lowerResMvCosts of that list and distance): bw4 * bh4
 * (x, y) pairs and costs, bw4 / bh4 = ((width / 4) + 7) >> 3, ((height / 4) + 7) >> 3.  The main encoder never reads them; the
 * parity tests do */
int  x265cu_fetch_hme_mvs(x265cu_ctx* ctx, int32_t slot, int32_t store, int32_t* mv_xy, int32_t* mv_costs);

/* ---- the host mirror of a decided frame in ONE asynchronous request: everything the main encoder reads of the frame's
 * Lowres (SURVEY 8b "output contract") goes to the caller's buffers on a dedicated copy stream, behind the work that
 * produces it (the frame's pre-lookahead, the batches that wrote the stores, the cuTree passes enqueued so far).  The call
 * returns at once; x265cu_mirror_wait blocks until that request has landed.  Destinations should be page-locked
 * (x265cu_pin_host) -- pageable ones make the enqueue itself wait for the copies.  At most X265CU_MIRROR_RING requests are in
 * flight; an older one is waited for when the ring wraps. */
#define X265CU_MIRROR_MAX_MV 36
#define X265CU_MIRROR_RING   8
typedef struct
{
    int32_t*  intra_cost;        /* ncu, or NULL */
    double*   qp_aq_offset;      /* ncu_full */
    double*   qp_cutree_offset;  /* ncu_full */
    int32_t*  inv_qscale_factor; /* ncu_full */
    void*     planes;            /* 4 * stride * plane_lines samples (de-tiled, Lowres::buffer[0]) */
    int32_t   n_mv;              /* MV stores to mirror, <= X265CU_MIRROR_MAX_MV */
    int32_t   mv_store[X265CU_MIRROR_MAX_MV];
    int32_t*  mv_dst[X265CU_MIRROR_MAX_MV];      /* ncu (x, y) pairs each: Lowres::lowresMvs[list][dist] */
    int32_t   cost_store;        /* cost store to mirror, -1 = none, 0 = the intra estimate */
    uint16_t* lowres_costs;      /* ncu */
    int32_t*  row_satds;         /* bh */
} x265cu_mirror_request;
int  x265cu_mirror_enqueue(x265cu_ctx* ctx, int32_t slot, const x265cu_mirror_request* req, int64_t* ticket);
int  x265cu_mirror_wait(x265cu_ctx* ctx, int64_t ticket);

/* frameCostRecalculate (slicetype.cpp:3802-3879) ahead of time: enqueued when the frame's qp offsets are final (the end of the
 * decision that outputs it), collected by x265cu_cost_recalc_get when RateControl asks -- by then it has long finished, so
 * getEstimatedPictureCost does not stall behind the cuTree work of later decisions.  The recalculated row sums stay in a
 * per-slot scratch until _get, which moves them into the cost store's rowSatds (where the reference's call leaves them). */
int  x265cu_cost_recalc_enqueue(x265cu_ctx* ctx, int32_t slot, int32_t cost_store, int32_t use_cutree_offsets);
/* synchronises with that one request; returns X265CU_ERR_BAD_ARG when nothing was enqueued for (slot, cost_store) */
int  x265cu_cost_recalc_get(x265cu_ctx* ctx, int32_t slot, int32_t cost_store, int64_t* score, int32_t* row_satds);

/* wait for everything enqueued on this context */
int  x265cu_sync(x265cu_ctx* ctx);

/* device-side stopwatch: CUDA events on the engine's compute stream (start ... stop -> elapsed ms) */
int  x265cu_timer_start(x265cu_ctx* ctx);
int  x265cu_timer_stop(x265cu_ctx* ctx, double* ms);

/* counters for bench.py: kernels launched and bytes copied since create; search_jobs / cost_jobs count the jobs that
 * ran (conditional jobs whose condition was false are not counted).  synchronises */
typedef struct { uint64_t kernel_launches, h2d_bytes, d2h_bytes, search_jobs, cost_jobs; } x265cu_counters;
int  x265cu_get_counters(x265cu_ctx* ctx, x265cu_counters* out);

/* timing hook for bench.py: device time (ms, CUDA events on the engine's stream) spent in each
 * kernel family since the last reset; enabling it adds two non-blocking event records per launch */
#define X265CU_K_LOWRES 0
#define X265CU_K_AQ     1
#define X265CU_K_INTRA  2
#define X265CU_K_SEARCH 3
#define X265CU_K_COST   4
#define X265CU_K_WEIGHT 5
#define X265CU_K_CUTREE 6
#define X265CU_K_COUNT  7
int  x265cu_profile_enable(x265cu_ctx* ctx, int32_t on);
int  x265cu_profile_get(x265cu_ctx* ctx, double ms[X265CU_K_COUNT], uint64_t launches[X265CU_K_COUNT], int32_t reset);
/* like x265cu_profile_get, but the time during which at least one kernel of the family was running (batches overlap,
 * so the sum of the launch durations can exceed the wall clock); call before a resetting x265cu_profile_get */
int  x265cu_profile_get_busy(x265cu_ctx* ctx, double busy_ms[X265CU_K_COUNT]);

#ifdef __cplusplus
}
#endif
#endif
