/* simengine.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Implements the engine C ABI (include/x265cu.h) on the CPU with the oracle's block-level
 * functions (oracle/la_oracle.c), so that the PRODUCT's host decision logic
 * (x265-amod_b200/host/lookahead.cpp) can be exercised against the real reference on machines
 * without a GPU.  It is linked only into tests/_build/libx265la_sim{8,10}.so by tests/build_sim.py;
 * the shipped libraries (libx265cu.so, libx265la.so) never contain or load it, and the shipped
 * engine has no CPU path at all.  Results produced through this file prove nothing about the
 * CUDA kernels -- those are checked by the `-m gpu` tests.
 */
#include "x265cu.h"
#include "la_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <vector>
/* the PRODUCT's generic search source (dia / hex / umh + subpel), compiled for the CPU with a scalar evaluator: with
 * X265SIM_GENERIC_ME=1 in the environment the --hme searches below run through it instead of the oracle's restatement, so
 * the text the CUDA kernel compiles is itself checked against the reference (control flow and arithmetic; the warp-level
 * evaluators are checked on the GPU) */
#include "../../x265-amod_b200/csrc/la_me_generic.cuh"

namespace {
struct CpuMeCtx
{
    or_pixel fenc[64];
    const or_pixel* ref[4];
    int stride;
    const uint16_t* mvcost;
    int mvpx, mvpy;
    int mvc(int qx, int qy) const { return (uint16_t)(mvcost[qx - mvpx] + mvcost[qy - mvpy]); }
    int sadFpel(int x, int y) const { return or_sad8x8(fenc, 8, ref[0] + x + (int64_t)y * stride, stride); }
    const or_pixel* mc(int qx, int qy, or_pixel* buf, int* st) const      /* lowres.h:71-96 */
    {
        const int hA = (qy & 2) | ((qx & 2) >> 1);
        const or_pixel* a = ref[hA] + (qx >> 2) + (int64_t)(qy >> 2) * stride;
        if (!((qx | qy) & 1)) { *st = stride; return a; }
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const or_pixel* b = ref[hB] + (qx2 >> 2) + (int64_t)(qy2 >> 2) * stride;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                buf[y * 8 + x] = (or_pixel)((a[y * stride + x] + b[y * stride + x] + 1) >> 1);
        *st = 8;
        return buf;
    }
    int qpelSad(int qx, int qy) const { or_pixel buf[64]; int st; const or_pixel* p = mc(qx, qy, buf, &st); return or_sad8x8(fenc, 8, p, st); }
    int qpelSatd(int qx, int qy) const { or_pixel buf[64]; int st; const or_pixel* p = mc(qx, qy, buf, &st); return or_satd8x8(fenc, 8, p, st); }
};

/* same contract as or_search_list_hme (oracle/la_oracle.h), the per-block search done by la::motionEstimateG */
void genericSearchList(const or_geom* g, int level, const or_pixel* fencPlane0, const or_pixel* const refPlanes[4],
                       const uint16_t* mvcost, int bBidir, int method, int merange, const int32_t* hmeMvs, const int32_t* hmeMvCosts,
                       int32_t* mvs, int32_t* mvCosts, int32_t* skipCount)
{
    const bool hme = level == 0;
    const int bw = hme ? g->bw4 : g->bw, bh = hme ? g->bh4 : g->bh, stride = hme ? g->stride4 : g->stride;
    int skips = 0;
    CpuMeCtx m;
    m.stride = stride; m.mvcost = mvcost;
    for (int cuY = bh - 1; cuY >= 0; cuY--)
        for (int cuX = bw - 1; cuX >= 0; cuX--)
        {
            const int cu = cuX + cuY * bw;
            const int64_t pel = 8 * cuX + (int64_t)8 * cuY * stride;
            for (int y = 0; y < 8; y++) memcpy(m.fenc + 8 * y, fencPlane0 + pel + (int64_t)y * stride, 8 * sizeof(or_pixel));
            for (int i = 0; i < 4; i++) m.ref[i] = refPlanes[i] + pel;
            const la::MV2 mvmin = { -cuX * 8 - 8, -cuY * 8 - 8 }, mvmax = { (bw - cuX - 1) * 8 + 8, (bh - cuY - 1) * 8 + 8 };
            la::MV2 cand[5]; int numc = 0;
            const bool lastRow = cuY == bh - 1;
            if (cuX < bw - 1) { cand[numc].x = mvs[2 * (cu + 1)]; cand[numc].y = mvs[2 * (cu + 1) + 1]; numc++; }
            if (!lastRow)
            {
                cand[numc].x = mvs[2 * (cu + bw)]; cand[numc].y = mvs[2 * (cu + bw) + 1]; numc++;
                if (cuX > 0) { cand[numc].x = mvs[2 * (cu + bw - 1)]; cand[numc].y = mvs[2 * (cu + bw - 1) + 1]; numc++; }
                if (cuX < bw - 1) { cand[numc].x = mvs[2 * (cu + bw + 1)]; cand[numc].y = mvs[2 * (cu + bw + 1) + 1]; numc++; }
            }
            const int cu4 = (cuX / 2) + (cuY / 2) * bw / 2;
            if (!hme && hmeMvs && hmeMvCosts[cu4] > 0) { cand[numc].x = hmeMvs[2 * cu4] * 2; cand[numc].y = hmeMvs[2 * cu4 + 1] * 2; numc++; }
            la::MV2 mvp = { 0, 0 };
            int skipCost = 0x7fffffff, mvpcost = 1 << 28;
            for (int i = 0; i < numc; i++)
            {
                const int cost = m.qpelSatd(cand[i].x, cand[i].y);
                if (cost < mvpcost) { mvpcost = cost; mvp = cand[i]; }
                if (!(mvp.x | mvp.y) && bBidir) skipCost = cost;
            }
            la::MV2 best;
            int fencCost = la::motionEstimateG(m, mvmin, mvmax, mvp, merange, method, best);
            if (skipCost < 64 && skipCost < fencCost && bBidir) { fencCost = skipCost; best.x = best.y = 0; skips++; }
            mvs[2 * cu] = best.x; mvs[2 * cu + 1] = best.y;
            mvCosts[cu] = fencCost;
        }
    if (skipCount) *skipCount = skips;
}
} // namespace

struct SimSlot
{
    std::vector<or_pixel> y, u, v;
    std::vector<or_pixel> planes;
    std::vector<or_pixel> planes4;                              /* --hme: the four 1/16-resolution planes */
    std::vector<std::vector<int32_t> > mvs4, mvCosts4;          /* --hme: level-0 results per MV store */
    std::vector<int32_t> intraCost, invQ, invQ8, rowSatds00;
    std::vector<uint8_t> intraMode;
    std::vector<uint16_t> lowresCosts00, propagate;
    std::vector<double> qpAq, qpCuTree;
    std::vector<std::vector<int32_t> > mvs, mvCosts;          /* per MV store */
    std::vector<int32_t> skipFlag;
    std::vector<std::vector<uint16_t> > costs;                /* per cost store */
    std::vector<std::vector<int32_t> > rowSatds;
    std::vector<x265cu_cost_result> results;
    std::vector<x265cu_hist_stats> hist;                        /* 0 or 1 entries (--hist-scenecut) */
    x265cu_frame_stats stats;
    std::vector<int32_t> recalcRows; int64_t recalcScore; int recalcStore;
};

struct x265cu_ctx
{
    x265cu_config cfg;
    or_geom g;
    x265cu_geometry geom;
    std::vector<uint16_t> mvcost;
    std::vector<SimSlot> slots;
    std::vector<or_pixel> wbuf;
    x265cu_counters counters;
    char err[64];
    /* sharded stream (x265cu_shard_config): same protocol as the CUDA engine, host memory instead of HBM */
    int rank, nranks;
    x265cu_exchange_fn exchange; void* exchangeUser;
    std::vector<int> owner;
    struct Seg { void* ptr; size_t bytes; int root; };
    std::vector<Seg> segs;
    std::vector<char> xbuf[8];
};

static bool simMine(const x265cu_ctx* c, int slot) { return c->nranks <= 1 || c->owner[slot] == c->rank; }
static void simSeg(x265cu_ctx* c, void* p, size_t bytes, int slot)
{
    if (c->nranks <= 1) return;
    x265cu_ctx::Seg s = { p, bytes, c->owner[slot] };
    c->segs.push_back(s);
}
static int simExchange(x265cu_ctx* c)
{
    if (c->nranks <= 1 || c->segs.empty()) return 0;
    uint64_t total[8] = { 0 };
    for (size_t i = 0; i < c->segs.size(); i++) total[c->segs[i].root] += c->segs[i].bytes;
    void* bufs[8];
    for (int r = 0; r < 8; r++) { c->xbuf[r].resize(total[r] ? total[r] : 1); bufs[r] = &c->xbuf[r][0]; }
    size_t off[8] = { 0 };
    for (size_t i = 0; i < c->segs.size(); i++)
    {
        const x265cu_ctx::Seg& s = c->segs[i];
        if (s.root == c->rank) memcpy(&c->xbuf[s.root][off[s.root]], s.ptr, s.bytes);
        off[s.root] += s.bytes;
    }
    if (c->exchange(c->exchangeUser, bufs, total, c->nranks, NULL) != 0) return X265CU_ERR_CUDA;
    memset(off, 0, sizeof(off));
    for (size_t i = 0; i < c->segs.size(); i++)
    {
        const x265cu_ctx::Seg& s = c->segs[i];
        if (s.root != c->rank) memcpy(s.ptr, &c->xbuf[s.root][off[s.root]], s.bytes);
        off[s.root] += s.bytes;
    }
    c->segs.clear();
    return 0;
}

extern "C" {

int x265cu_device_count(void) { return 0; }
const char* x265cu_strerror(int s) { return s == 0 ? "ok" : "simengine error"; }
const char* x265cu_last_error(const x265cu_ctx*) { return ""; }

int x265cu_create(const x265cu_config* cfg, x265cu_ctx** out)
{
    if (cfg->depth != or_depth()) return X265CU_ERR_BAD_ARG;
    if (cfg->hist_stats && cfg->depth != 8) return X265CU_ERR_UNSUPPORTED;
    if (cfg->aq_mode > 3 && cfg->fade_stats) return X265CU_ERR_UNSUPPORTED;
    if (cfg->hme && (cfg->hme_search[0] < 0 || cfg->hme_search[0] > 5 || cfg->hme_search[0] == 4 || cfg->hme_search[1] < 0 || cfg->hme_search[1] > 5 || cfg->hme_search[1] == 4)) return X265CU_ERR_UNSUPPORTED;
    x265cu_ctx* c = new x265cu_ctx;
    c->cfg = *cfg;
    or_geom_init(&c->g, cfg->width, cfg->height, cfg->max_cu_size);
    c->g.rowsPerSlice = 0;      /* per search job, x265cu_search_job::sliced */
    c->g.qg8 = cfg->qg_size == 8;
    c->mvcost.assign(cfg->mvcost, cfg->mvcost + 2 * (size_t)cfg->mvcost_half + 1);
    memset(&c->geom, 0, sizeof(c->geom));
    c->geom.low_width = c->g.w; c->geom.low_height = c->g.h; c->geom.bw = c->g.bw; c->geom.bh = c->g.bh;
    c->geom.ncu = c->g.ncu; c->geom.stride = c->g.stride; c->geom.plane_lines = c->g.planeLines;
    c->geom.margin_x = c->g.mx; c->geom.margin_y = c->g.my; c->geom.nb = cfg->bframes + 2;
    c->geom.n_mv_stores = (cfg->mv_store_kinds > 0 ? cfg->mv_store_kinds : 3) * c->geom.nb;
    c->geom.n_cost_stores = (cfg->cost_variants > 0 ? cfg->cost_variants : 2) * c->geom.nb * c->geom.nb;
    c->geom.ncu_full = c->g.qg8 ? 4 * c->g.ncu : c->g.ncu;
    c->slots.resize(cfg->max_slots);
    c->rank = 0; c->nranks = 1; c->exchange = NULL; c->exchangeUser = NULL; c->owner.assign(cfg->max_slots, 0);
    memset(&c->counters, 0, sizeof(c->counters));
    *out = c;
    return 0;
}

void x265cu_destroy(x265cu_ctx* c) { delete c; }
int x265cu_get_geometry(const x265cu_ctx* c, x265cu_geometry* g) { *g = c->geom; return 0; }
int x265cu_sm_partition(const x265cu_ctx*, int32_t* a, int32_t* b) { if (a) *a = 0; if (b) *b = 0; return 0; }
int x265cu_pin_host(x265cu_ctx*, void*, uint64_t) { return 0; }
int x265cu_unpin_host(x265cu_ctx*, void*) { return 0; }
int x265cu_sync(x265cu_ctx*) { return 0; }
int x265cu_timer_start(x265cu_ctx*) { return 0; }
int x265cu_timer_stop(x265cu_ctx*, double* ms) { *ms = 0; return 0; }
int x265cu_get_counters(x265cu_ctx* c, x265cu_counters* o) { *o = c->counters; return 0; }
int x265cu_batch_begin(x265cu_ctx*, int64_t* id) { static int64_t n = 0; if (id) *id = n++; return 0; }
int x265cu_batch_end(x265cu_ctx*) { return 0; }
int x265cu_batches_in_flight(x265cu_ctx*) { static int n = 0; return n++ % 3; }    /* exercise every answer (0, 1, 2) */
int x265cu_shard_config(x265cu_ctx* c, int32_t rank, int32_t nranks, x265cu_exchange_fn fn, void* user)
{
    if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks || (nranks > 1 && !fn)) return X265CU_ERR_BAD_ARG;
    c->rank = rank; c->nranks = nranks; c->exchange = fn; c->exchangeUser = user;
    return 0;
}
int x265cu_slot_owner(x265cu_ctx* c, int32_t slot, int32_t owner) { c->owner[slot] = owner; return 0; }
int x265cu_frame_ready(x265cu_ctx*, int32_t slot)
{
    if (getenv("X265CU_FRAME_READY_NEVER")) return 0;    /* same test hook as the engine: every pair takes the assumed-weights path */
    return (slot % 3) != 1;     /* exercise both answers */
}
int x265cu_profile_get_busy(x265cu_ctx*, double* ms) { for (int i = 0; i < X265CU_K_COUNT; i++) ms[i] = 0; return 0; }
int x265cu_profile_enable(x265cu_ctx*, int32_t) { return 0; }
int x265cu_profile_get(x265cu_ctx*, double* ms, uint64_t* n, int32_t) { for (int i = 0; i < X265CU_K_COUNT; i++) { ms[i] = 0; n[i] = 0; } return 0; }

int x265cu_frame_upload(x265cu_ctx* c, int32_t slot, const void* y, const void* u, const void* v, int32_t sy, int32_t sc)
{
    SimSlot& s = c->slots[slot];
    const or_geom& g = c->g;
    if (c->cfg.hist_stats && (!u || !v)) return X265CU_ERR_BAD_ARG;     /* as the engine: the statistics read chroma */
    const int W = g.picW, H = g.picH, CW = (W + 1) / 2, CH = (H + 1) / 2;
    s.y.resize((size_t)W * H);
    for (int r = 0; r < H; r++) memcpy(&s.y[(size_t)r * W], (const or_pixel*)y + (size_t)r * sy, W * sizeof(or_pixel));
    if (u && v)
    {
        s.u.resize((size_t)CW * CH); s.v.resize((size_t)CW * CH);
        for (int r = 0; r < CH; r++)
        {
            memcpy(&s.u[(size_t)r * CW], (const or_pixel*)u + (size_t)r * sc, CW * sizeof(or_pixel));
            memcpy(&s.v[(size_t)r * CW], (const or_pixel*)v + (size_t)r * sc, CW * sizeof(or_pixel));
        }
    }
    else { s.u.clear(); s.v.clear(); }
    s.planes.assign((size_t)(4 * g.planeSize), 0);
    or_lowres_init(&g, &s.y[0], W, &s.planes[0]);
    if (c->cfg.hme)
    {
        s.planes4.assign((size_t)(4 * g.planeSize4), 0);
        or_lowerres_init(&g, &s.planes[g.padOffset], &s.planes4[0]);
    }
    const int ncu = g.ncu, nb = c->geom.nb;
    /* the qp-offset arrays are allocated zeroed once per Lowres in the reference (lowres.cpp:98-106) and entries the
     * running AQ index never reaches stay zero */
    const size_t nAq = (size_t)c->geom.ncu_full + 2 * g.bw + 2 + (g.qg8 ? ((g.picW + 7) / 8) * ((g.picH + 7) / 8) : 0);
    /* (with aq-mode 0 but weightp on, the arrays exist and nothing ever writes them: invQscaleFactor stays 0 and every AQ-scaled
     * cost is 0 in the reference, slicetype.cpp:487-511, 4226-4229) */
    s.intraCost.assign(ncu, 0); s.invQ.assign(nAq, 0); s.invQ8.assign(ncu, g.qg8 && c->cfg.aq_mode ? 256 : 0); s.intraMode.assign(ncu, 0);
    s.lowresCosts00.assign(ncu, 0); s.rowSatds00.assign(g.bh, 0); s.propagate.assign(ncu, 0);
    s.qpAq.assign(nAq, 0.0); s.qpCuTree.assign(nAq, 0.0);
    const int nmv = c->geom.n_mv_stores, ncs = c->geom.n_cost_stores;
    s.mvs.assign(nmv, std::vector<int32_t>()); s.mvCosts.assign(nmv, std::vector<int32_t>());
    s.skipFlag.assign(nmv, 0);
    s.mvs4.assign(nmv, std::vector<int32_t>()); s.mvCosts4.assign(nmv, std::vector<int32_t>());
    s.costs.assign(ncs, std::vector<uint16_t>()); s.rowSatds.assign(ncs, std::vector<int32_t>());
    s.results.assign(ncs, x265cu_cost_result());
    memset(&s.stats, 0, sizeof(s.stats));
    s.recalcStore = -1;
    if (c->cfg.need_aq)
        or_aq_frame(&g, &s.y[0], W, s.u.empty() ? NULL : &s.u[0], s.v.empty() ? NULL : &s.v[0], CW,
                    c->cfg.aq_mode, c->cfg.aq_strength, c->cfg.need_wp_stats, &s.qpAq[0], &s.qpCuTree[0], &s.invQ[0],
                    NULL, s.stats.wp_ssd, s.stats.wp_sum);
    if (c->cfg.need_aq && c->cfg.fade_stats)
        s.stats.frame_variance = or_fade_variance(&g, &s.y[0], W, s.u.empty() ? NULL : &s.u[0], s.v.empty() ? NULL : &s.v[0], CW,
                                                  c->cfg.need_wp_stats, s.stats.wp_ssd, s.stats.wp_sum);
    if (c->cfg.hist_stats)
    {
        static_assert(sizeof(or_hist_stats_t) == sizeof(x265cu_hist_stats), "same layout");
        s.hist.resize(1);
        or_hist_stats(&g, &s.y[0], W, s.u.empty() ? NULL : &s.u[0], s.v.empty() ? NULL : &s.v[0], CW, &s.planes[g.padOffset],
                      (or_hist_stats_t*)&s.hist[0]);
    }
    if (c->cfg.need_aq && g.qg8) or_invq8x8(&g, &s.invQ[0], &s.invQ8[0]);
    or_intra_estimate(&g, &s.planes[g.padOffset], c->cfg.need_aq ? (g.qg8 ? &s.invQ8[0] : &s.invQ[0]) : NULL, &s.intraCost[0], &s.intraMode[0],
                      &s.lowresCosts00[0], &s.rowSatds00[0], &s.stats.cost_est, &s.stats.cost_est_aq);
    c->counters.h2d_bytes += (uint64_t)W * H * sizeof(or_pixel) * 3 / 2;
    return 0;
}

int x265cu_frame_hist_get(x265cu_ctx* c, int32_t slot, x265cu_hist_stats* out)
{
    if (!c->cfg.hist_stats || c->slots[slot].hist.empty()) return X265CU_ERR_BAD_ARG;
    *out = c->slots[slot].hist[0];
    return 0;
}

int x265cu_frame_stats_get(x265cu_ctx* c, const int32_t* slots, int32_t n, x265cu_frame_stats* out)
{
    for (int i = 0; i < n; i++) out[i] = c->slots[slots[i]].stats;
    return 0;
}

static void planePtrs(const x265cu_ctx* c, const std::vector<or_pixel>& buf, const or_pixel* p[4])
{
    for (int i = 0; i < 4; i++) p[i] = &buf[(size_t)(i * c->g.planeSize + c->g.padOffset)];
}

int x265cu_search_batch(x265cu_ctx* c, const x265cu_search_job* jobs, int32_t n)
{
    const or_geom& g = c->g;
    for (int i = 0; i < n; i++)
    {
        const x265cu_search_job& j = jobs[i];
        SimSlot& f = c->slots[j.fenc_slot]; SimSlot& r = c->slots[j.ref_slot];
        f.mvs[j.store].resize(2 * g.ncu); f.mvCosts[j.store].resize(g.ncu);
        simSeg(c, &f.mvs[j.store][0], 2 * g.ncu * sizeof(int32_t), j.fenc_slot);
        simSeg(c, &f.mvCosts[j.store][0], g.ncu * sizeof(int32_t), j.fenc_slot);
        simSeg(c, &f.skipFlag[j.store], sizeof(int32_t), j.fenc_slot);
        if (!simMine(c, j.fenc_slot)) continue;
        if (j.cond_store >= 0 && !f.skipFlag[j.cond_store]) continue;     /* conditional job, x265cu.h */
        c->counters.search_jobs++;
        const or_pixel* rp[4];
        if (j.weighted)
        {
            c->wbuf.resize((size_t)(4 * g.planeSize));
            or_weight_planes(&g, &r.planes[0], &c->wbuf[0], 4, j.w_scale, j.w_denom, j.w_offset);
            planePtrs(c, c->wbuf, rp);
        }
        else
            planePtrs(c, r.planes, rp);
        f.mvs[j.store].assign(2 * g.ncu, 0); f.mvCosts[j.store].assign(g.ncu, 0);
        or_geom gj = g;
        gj.rowsPerSlice = j.sliced ? c->cfg.rows_per_slice : 0;
        if (c->cfg.hme)
        {
            /* level 0 on the 1/16-resolution planes (never weighted), then level 1 fed with its vectors; a skip at either
             * level makes the B-context search differ from the P-context one (x265cu_search_job::cond_store) */
            const int n4 = g.bw4 * g.bh4;
            f.mvs4[j.store].assign(2 * n4, 0); f.mvCosts4[j.store].assign(n4, 0);
            const or_pixel* rp4[4];
            for (int k = 0; k < 4; k++) rp4[k] = &r.planes4[(size_t)(k * g.planeSize4 + g.padOffset4)];
            int32_t skips0 = 0, skips1 = 0;
            static const bool generic = getenv("X265SIM_GENERIC_ME") && atoi(getenv("X265SIM_GENERIC_ME"));
            void (*searchList)(const or_geom*, int, const or_pixel*, const or_pixel* const*, const uint16_t*, int, int, int, const int32_t*,
                               const int32_t*, int32_t*, int32_t*, int32_t*) = generic ? genericSearchList : or_search_list_hme;
            searchList(&gj, 0, &f.planes4[g.padOffset4], rp4, &c->mvcost[c->cfg.mvcost_half], j.bidir_ctx,
                               c->cfg.hme_search[0], c->cfg.hme_range[0], NULL, NULL, &f.mvs4[j.store][0], &f.mvCosts4[j.store][0], &skips0);
            searchList(&gj, 1, &f.planes[g.padOffset], rp, &c->mvcost[c->cfg.mvcost_half], j.bidir_ctx,
                               c->cfg.hme_search[1], c->cfg.hme_range[1], &f.mvs4[j.store][0], &f.mvCosts4[j.store][0],
                               &f.mvs[j.store][0], &f.mvCosts[j.store][0], &skips1);
            f.skipFlag[j.store] = skips0 + skips1;
        }
        else
        or_search_list(&gj, &f.planes[g.padOffset], rp, &c->mvcost[c->cfg.mvcost_half], j.bidir_ctx,
                       &f.mvs[j.store][0], &f.mvCosts[j.store][0], &f.skipFlag[j.store]);
        c->counters.kernel_launches++;
    }
    return simExchange(c);
}

int x265cu_search_flags_get(x265cu_ctx* c, const int32_t* slots, const int32_t* stores, int32_t n, int32_t* flags)
{
    for (int i = 0; i < n; i++) flags[i] = c->slots[slots[i]].skipFlag[stores[i]] != 0;
    return 0;
}

int x265cu_cost_batch(x265cu_ctx* c, const x265cu_cost_job* jobs, int32_t n)
{
    const or_geom& g = c->g;
    for (int i = 0; i < n; i++)
    {
        const x265cu_cost_job& j = jobs[i];
        SimSlot& b = c->slots[j.b_slot]; SimSlot& p0 = c->slots[j.p0_slot]; SimSlot& p1 = c->slots[j.p1_slot];
        b.costs[j.out].resize(g.ncu); b.rowSatds[j.out].resize(g.bh);
        simSeg(c, &b.costs[j.out][0], g.ncu * sizeof(uint16_t), j.b_slot);
        simSeg(c, &b.rowSatds[j.out][0], g.bh * sizeof(int32_t), j.b_slot);
        simSeg(c, &b.results[j.out], sizeof(x265cu_cost_result), j.b_slot);
        if (!simMine(c, j.b_slot)) continue;
        if (j.cond_store >= 0 && !b.skipFlag[j.cond_store]) continue;
        c->counters.cost_jobs++;
        const or_pixel *r0[4], *r1[4];
        planePtrs(c, p0.planes, r0); planePtrs(c, p1.planes, r1);
        const bool bidir = j.l1_store >= 0;
        b.costs[j.out].assign(g.ncu, 0); b.rowSatds[j.out].assign(g.bh, 0);
        x265cu_cost_result& res = b.results[j.out];
        or_frame_cost(&g, &b.planes[g.padOffset], r0, bidir ? r1 : NULL,
                      &b.mvs[j.l0_store][0], &b.mvCosts[j.l0_store][0],
                      bidir ? &b.mvs[j.l1_store][0] : NULL, bidir ? &b.mvCosts[j.l1_store][0] : NULL,
                      &b.intraCost[0], c->cfg.need_aq ? (g.qg8 ? &b.invQ8[0] : &b.invQ[0]) : NULL, &b.costs[j.out][0], &b.rowSatds[j.out][0],
                      &res.cost_est, &res.cost_est_aq, &res.intra_mbs);
        c->counters.kernel_launches++;
    }
    return simExchange(c);
}

int x265cu_cost_results_get(x265cu_ctx* c, const int32_t* slots, const int32_t* outs, int32_t n, x265cu_cost_result* res)
{
    for (int i = 0; i < n; i++) res[i] = c->slots[slots[i]].results[outs[i]];
    return 0;
}

int x265cu_weight_cost_batch(x265cu_ctx* c, const x265cu_wcost_job* jobs, int32_t n, uint32_t* costs)
{
    const or_geom& g = c->g;
    for (int i = 0; i < n; i++)
    {
        const x265cu_wcost_job& j = jobs[i];
        SimSlot& f = c->slots[j.fenc_slot]; SimSlot& r = c->slots[j.ref_slot];
        const or_pixel* ref0 = &r.planes[g.padOffset];
        if (j.weighted)
        {
            c->wbuf.resize((size_t)(4 * g.planeSize));
            or_weight_planes(&g, &r.planes[0], &c->wbuf[0], 1, j.w_scale, j.w_denom, j.w_offset);
            ref0 = &c->wbuf[g.padOffset];
        }
        costs[i] = or_weight_cost_luma(&g, &f.planes[g.padOffset], ref0, &f.intraCost[0]);
    }
    return 0;
}

int x265cu_cutree_reset(x265cu_ctx* c, int32_t slot)
{
    std::fill(c->slots[slot].propagate.begin(), c->slots[slot].propagate.end(), 0);
    return 0;
}

int x265cu_cutree_propagate(x265cu_ctx* c, int32_t bs, int32_t p0s, int32_t p1s, int32_t cost_store, int32_t l0, int32_t l1,
                            int32_t referenced, int32_t bipred_weight, double fps_factor)
{
    SimSlot& b = c->slots[bs];
    or_cutree_propagate(&c->g, &b.intraCost[0], &b.costs[cost_store][0], c->g.qg8 ? &b.invQ8[0] : &b.invQ[0], &b.mvs[l0][0],
                        l1 >= 0 ? &b.mvs[l1][0] : &b.mvs[l0][0], &b.propagate[0],
                        &c->slots[p0s].propagate[0], &c->slots[p1s].propagate[0], referenced, bipred_weight, fps_factor);
    return 0;
}

int x265cu_cutree_finish(x265cu_ctx* c, int32_t slot, int32_t fps_fix8, double weightdelta, double strength)
{
    SimSlot& s = c->slots[slot];
    or_cutree_finish(&c->g, &s.intraCost[0], c->g.qg8 ? &s.invQ8[0] : &s.invQ[0], &s.propagate[0], &s.qpAq[0], &s.qpCuTree[0], fps_fix8, weightdelta, strength);
    return 0;
}

int x265cu_cost_recalc(x265cu_ctx* c, int32_t slot, int32_t cost_store, int32_t use_cutree, int64_t* score, int32_t* rows)
{
    SimSlot& s = c->slots[slot];
    std::vector<uint16_t>& costs = cost_store == 0 ? s.lowresCosts00 : s.costs[cost_store];
    std::vector<int32_t>& rs = cost_store == 0 ? s.rowSatds00 : s.rowSatds[cost_store];
    *score = or_frame_cost_recalc(&c->g, &costs[0], use_cutree ? &s.qpCuTree[0] : &s.qpAq[0], &rs[0]);
    if (rows) memcpy(rows, &rs[0], c->g.bh * sizeof(int32_t));
    return 0;
}

int x265cu_vbv_row_costs(x265cu_ctx* c, int32_t slot, int32_t cost_store, int32_t qp_source, int32_t ctu_rows_lowres,
                         int32_t pir_start, int32_t pir_end, int32_t n_rows, uint32_t* satd, uint32_t* intra,
                         uint16_t* cost_for_rc, int32_t* intra_scaled)
{
    SimSlot& s = c->slots[slot];
    const std::vector<uint16_t>& costs = cost_store == 0 ? s.lowresCosts00 : s.costs[cost_store];
    std::vector<uint32_t> a(n_rows), b(n_rows);
    std::vector<uint16_t> rc(c->g.ncu); std::vector<int32_t> ic(c->g.ncu);
    const double* qp = (qp_source == 0 || !c->cfg.need_aq) ? NULL : (qp_source == 2 ? &s.qpCuTree[0] : &s.qpAq[0]);
    or_vbv_rows(&c->g, &costs[0], &s.intraCost[0], qp, ctu_rows_lowres, pir_start, pir_end, n_rows, &a[0], &b[0], &rc[0], &ic[0]);
    if (satd) memcpy(satd, &a[0], n_rows * 4);
    if (intra) memcpy(intra, &b[0], n_rows * 4);
    if (cost_for_rc) memcpy(cost_for_rc, &rc[0], c->g.ncu * 2);
    if (intra_scaled) memcpy(intra_scaled, &ic[0], c->g.ncu * 4);
    return 0;
}

int x265cu_fetch_mvs(x265cu_ctx* c, int32_t slot, int32_t store, int32_t* mv, int32_t* cost);
int x265cu_fetch_costs(x265cu_ctx* c, int32_t slot, int32_t store, uint16_t* costs, int32_t* rows);
int x265cu_fetch_frame(x265cu_ctx* c, int32_t slot, const x265cu_frame_out* o);

/* the asynchronous mirror, done on the spot */
int x265cu_mirror_enqueue(x265cu_ctx* c, int32_t slot, const x265cu_mirror_request* q, int64_t* ticket)
{
    x265cu_frame_out o;
    memset(&o, 0, sizeof(o));
    o.intra_cost = q->intra_cost; o.qp_aq_offset = q->qp_aq_offset; o.qp_cutree_offset = q->qp_cutree_offset;
    o.inv_qscale_factor = q->inv_qscale_factor; o.planes = q->planes;
    if (q->cost_store == 0) { o.lowres_costs00 = q->lowres_costs; o.row_satds00 = q->row_satds; }
    x265cu_fetch_frame(c, slot, &o);
    for (int i = 0; i < q->n_mv; i++) x265cu_fetch_mvs(c, slot, q->mv_store[i], q->mv_dst[i], NULL);
    if (q->cost_store >= 2) x265cu_fetch_costs(c, slot, q->cost_store, q->lowres_costs, q->row_satds);
    static int64_t n = 0;
    if (ticket) *ticket = n++;
    return 0;
}
int x265cu_mirror_wait(x265cu_ctx*, int64_t) { return 0; }

int x265cu_cost_recalc(x265cu_ctx* c, int32_t slot, int32_t cost_store, int32_t use_cutree, int64_t* score, int32_t* rows);
/* ahead-of-time recalculation: computed at enqueue into a side buffer, published into the store's rowSatds at get */
int x265cu_cost_recalc_enqueue(x265cu_ctx* c, int32_t slot, int32_t cost_store, int32_t use_cutree)
{
    SimSlot& s = c->slots[slot];
    std::vector<int32_t>& rs = cost_store == 0 ? s.rowSatds00 : s.rowSatds[cost_store];
    std::vector<int32_t> keep(rs);
    s.recalcRows.resize(c->g.bh);
    x265cu_cost_recalc(c, slot, cost_store, use_cutree, &s.recalcScore, &s.recalcRows[0]);
    rs = keep;
    s.recalcStore = cost_store;
    return 0;
}
int x265cu_cost_recalc_get(x265cu_ctx* c, int32_t slot, int32_t cost_store, int64_t* score, int32_t* rows)
{
    SimSlot& s = c->slots[slot];
    if (s.recalcStore != cost_store) return X265CU_ERR_BAD_ARG;
    *score = s.recalcScore;
    if (rows) memcpy(rows, &s.recalcRows[0], c->g.bh * sizeof(int32_t));
    (cost_store == 0 ? s.rowSatds00 : s.rowSatds[cost_store]) = s.recalcRows;
    s.recalcStore = -1;
    return 0;
}

int x265cu_fetch_frame(x265cu_ctx* c, int32_t slot, const x265cu_frame_out* o)
{
    SimSlot& s = c->slots[slot];
    const int ncu = c->g.ncu, nfull = c->geom.ncu_full;
    if (o->intra_cost) memcpy(o->intra_cost, &s.intraCost[0], ncu * 4);
    if (o->intra_mode) memcpy(o->intra_mode, &s.intraMode[0], ncu);
    if (o->qp_aq_offset) memcpy(o->qp_aq_offset, &s.qpAq[0], nfull * 8);
    if (o->qp_cutree_offset) memcpy(o->qp_cutree_offset, &s.qpCuTree[0], nfull * 8);
    if (o->inv_qscale_factor)
    {
        if (c->cfg.need_aq) memcpy(o->inv_qscale_factor, &s.invQ[0], nfull * 4);
        else for (int i = 0; i < nfull; i++) o->inv_qscale_factor[i] = 256;
    }
    if (o->propagate_cost) memcpy(o->propagate_cost, &s.propagate[0], ncu * 2);
    if (o->planes) memcpy(o->planes, &s.planes[0], (size_t)(4 * c->g.planeSize) * sizeof(or_pixel));
    if (o->lowres_costs00) memcpy(o->lowres_costs00, &s.lowresCosts00[0], ncu * 2);
    if (o->row_satds00) memcpy(o->row_satds00, &s.rowSatds00[0], c->g.bh * 4);
    return 0;
}

int x265cu_fetch_mvs(x265cu_ctx* c, int32_t slot, int32_t store, int32_t* mv, int32_t* cost)
{
    SimSlot& s = c->slots[slot];
    if (s.mvs[store].empty()) return X265CU_ERR_BAD_ARG;
    if (mv) memcpy(mv, &s.mvs[store][0], c->g.ncu * 8);
    if (cost) memcpy(cost, &s.mvCosts[store][0], c->g.ncu * 4);
    return 0;
}

int x265cu_fetch_hme_mvs(x265cu_ctx* c, int32_t slot, int32_t store, int32_t* mv, int32_t* cost)
{
    SimSlot& s = c->slots[slot];
    if (!c->cfg.hme || s.mvs4[store].empty()) return X265CU_ERR_BAD_ARG;
    const int n4 = c->g.bw4 * c->g.bh4;
    if (mv) memcpy(mv, &s.mvs4[store][0], n4 * 8);
    if (cost) memcpy(cost, &s.mvCosts4[store][0], n4 * 4);
    return 0;
}

int x265cu_fetch_costs(x265cu_ctx* c, int32_t slot, int32_t store, uint16_t* costs, int32_t* rows)
{
    SimSlot& s = c->slots[slot];
    if (s.costs[store].empty()) return X265CU_ERR_BAD_ARG;
    if (costs) memcpy(costs, &s.costs[store][0], c->g.ncu * 2);
    if (rows) memcpy(rows, &s.rowSatds[store][0], c->g.bh * 4);
    return 0;
}

} // extern "C"
