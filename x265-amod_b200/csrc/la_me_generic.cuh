/* la_me_generic.cuh -- the integer motion searches --hme selects per level (dia / hex / umh) + the lowres subpel refinement,
 * star and full; sea is not built), written against a small evaluator interface instead of the warp-wide lockstep of the default search (la_kernels.cuh
 * motionEstimate, which stays as it is: HEX over 16 is all the lookahead runs without --hme).
 *
 * Reference semantics: MotionEstimate::motionEstimate with numCandidates == 0, subpelRefine 1, a lowres reference
 * (source/encoder/motion.cpp:764-868 prologue + DIA, :870-969 HEX + square refine, :971-1160 UMH, :387-630 + :1157-1264 STAR,
 * :1473-1528 subpel).
 *
 * `Ctx` provides
 *     int  sadFpel(int x, int y)          SAD of the source block against the reference block displaced by (x, y) full pels
 *     int  qpelSad / qpelSatd(int qx, int qy)   ReferencePlanes::lowresQPelCost at a quarter-pel vector (lowres.h:98-124)
 *     int  mvc(int qx, int qy)            BitCost::mvcost relative to (mvpx, mvpy) (bitcost.h:46)
 *     int  mvpx, mvpy                     written here (setMVP)
 * On the GPU one 8-lane group evaluates a candidate (MeCtxG, la_kernels.cuh) and the four groups of a warp run this
 * function independently (their shuffles name only their own lanes).  The file has no CUDA-only constructs, so the CPU
 * suite compiles the very same text with a scalar evaluator and checks it against the reference's results
 * (tests/test_oracle_and_host.py::test_generic_search_source_on_cpu).
 */
#pragma once

#ifdef __CUDACC__
#define LA_HD __device__ __forceinline__
#define LA_CONST_TABLE __device__ const
#else
#define LA_HD static inline
#define LA_CONST_TABLE static const
#endif

namespace la {

struct MV2 { int x, y; };

enum { LA_DIA_SEARCH = 0, LA_HEX_SEARCH = 1, LA_UMH_SEARCH = 2, LA_STAR_SEARCH = 3, LA_FULL_SEARCH = 5 };     /* X265_*_SEARCH, x265.h (4 = SEA: not built) */

LA_CONST_TABLE signed char g_hex2[8][2] = { {-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0} };    /* motion.cpp:64 */
LA_CONST_TABLE unsigned char g_mod6m1[8] = { 5, 0, 1, 2, 3, 4, 5, 0 };                                                 /* :65 */
LA_CONST_TABLE signed char g_square1[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1} };   /* :66 */
LA_CONST_TABLE signed char g_hex4[16][2] = { {0, -4}, {0, 4}, {-2, -3}, {2, -3}, {-4, -2}, {4, -2}, {-4, -1}, {4, -1},
                                             {-4, 0}, {4, 0}, {-4, 1}, {4, 1}, {-4, 2}, {4, 2}, {-2, 3}, {2, 3} };      /* :67-73 */

/* the two outer neighbours of each of the 8 points around the centre (motion.cpp:74-84) */
LA_CONST_TABLE signed char g_starOffsets[16][2] = { {-1, 0}, {0, -1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {-1, -1},
                                                    {1, -1}, {1, 1}, {-1, 0}, {0, 1}, {-1, 1}, {1, 1}, {1, 0}, {0, 1} };

LA_HD int gmin(int a, int b) { return a < b ? a : b; }
LA_HD int gmax(int a, int b) { return a > b ? a : b; }
LA_HD bool gInRange(MV2 v, MV2 lo, MV2 hi) { return v.x >= lo.x && v.x <= hi.x && v.y >= lo.y && v.y <= hi.y; }     /* MV::checkRange */

/* one full-pel candidate: COST_MV (motion.cpp:263-269) */
template <typename Ctx>
LA_HD void gCostMv(Ctx& m, int mx, int my, int& bcost, MV2& bmv)
{
    const int cost = m.sadFpel(mx, my) + m.mvc(mx << 2, my << 2);
    if (cost < bcost) { bcost = cost; bmv.x = mx; bmv.y = my; }
}

/* one of the four candidates of COST_MV_X4 (motion.cpp:309-330): measured around omv; only its y is range-checked */
template <typename Ctx>
LA_HD void gCostMvO(Ctx& m, MV2 omv, int dx, int dy, MV2 mvmin, MV2 mvmax, int& bcost, MV2& bmv)
{
    const int cost = m.sadFpel(omv.x + dx, omv.y + dy) + m.mvc((omv.x + dx) << 2, (omv.y + dy) << 2);
    if ((omv.y + dy >= mvmin.y) & (omv.y + dy <= mvmax.y))
        if (cost < bcost) { bcost = cost; bmv.x = omv.x + dx; bmv.y = omv.y + dy; }
}

template <typename Ctx>
LA_HD void gCostMvX4(Ctx& m, MV2 omv, int x0, int y0, int x1, int y1, int x2, int y2, int x3, int y3, MV2 mvmin, MV2 mvmax,
                     int& bcost, MV2& bmv)
{
    gCostMvO(m, omv, x0, y0, mvmin, mvmax, bcost, bmv);
    gCostMvO(m, omv, x1, y1, mvmin, mvmax, bcost, bmv);
    gCostMvO(m, omv, x2, y2, mvmin, mvmax, bcost, bmv);
    gCostMvO(m, omv, x3, y3, mvmin, mvmax, bcost, bmv);
}

/* CROSS (motion.cpp:356-385) */
template <typename Ctx>
LA_HD void gCross(Ctx& m, MV2 omv, int start, int xMax, int yMax, MV2 mvmin, MV2 mvmax, int& bcost, MV2& bmv)
{
    int i = start;
    if (xMax <= gmin(mvmax.x - omv.x, omv.x - mvmin.x))
        for (; i < xMax - 2; i += 4)
            gCostMvX4(m, omv, i, 0, -i, 0, i + 2, 0, -i - 2, 0, mvmin, mvmax, bcost, bmv);
    for (; i < xMax; i += 2)
    {
        if (omv.x + i <= mvmax.x) gCostMv(m, omv.x + i, omv.y, bcost, bmv);
        if (omv.x - i >= mvmin.x) gCostMv(m, omv.x - i, omv.y, bcost, bmv);
    }
    i = start;
    if (yMax <= gmin(mvmax.y - omv.y, omv.y - mvmin.y))
        for (; i < yMax - 2; i += 4)
            gCostMvX4(m, omv, 0, i, 0, -i, 0, i + 2, 0, -i - 2, mvmin, mvmax, bcost, bmv);
    for (; i < yMax; i += 2)
    {
        if (omv.y + i <= mvmax.y) gCostMv(m, omv.x, omv.y + i, bcost, bmv);
        if (omv.y - i >= mvmin.y) gCostMv(m, omv.x, omv.y - i, bcost, bmv);
    }
}

/* the front half of UMH (motion.cpp:971-1160, numCandidates == 0); returns true when the search goes on to the hexagon
 * refinement (`goto me_hex2`), false when it stopped early or left the range */
template <typename Ctx>
LA_HD bool gUmh(Ctx& m, MV2 mvmin, MV2 mvmax, MV2 pmv /* full-pel */, int merange, int& bcost, MV2& bmv)
{
    MV2 omv = bmv;
    int crossStart = 1;
    const int ucost1 = bcost;
    omv = pmv;                                                              /* DIA1_ITER(pmv.x, pmv.y) */
    gCostMvX4(m, omv, 0, -1, 0, 1, -1, 0, 1, 0, mvmin, mvmax, bcost, bmv);
    if (pmv.x | pmv.y)
    {
        omv.x = 0; omv.y = 0;
        gCostMvX4(m, omv, 0, -1, 0, 1, -1, 0, 1, 0, mvmin, mvmax, bcost, bmv);
    }
    const int ucost2 = bcost;
    if ((bmv.x | bmv.y) && (bmv.x != pmv.x || bmv.y != pmv.y))
    {
        omv = bmv;
        gCostMvX4(m, omv, 0, -1, 0, 1, -1, 0, 1, 0, mvmin, mvmax, bcost, bmv);
    }
    if (bcost == ucost2)
        crossStart = 3;
    omv = bmv;
    /* SAD_THRESH(v) = bcost < (v >> 4) * sizeScale[LUMA_8x8], sizeScale = (8 * 8) >> 4 (motion.cpp:61,126) */
    if (bcost == ucost2 && bcost < (2000 >> 4) * 4)
    {
        gCostMvX4(m, omv, 0, -2, -1, -1, 1, -1, -2, 0, mvmin, mvmax, bcost, bmv);
        gCostMvX4(m, omv, 2, 0, -1, 1, 1, 1, 0, 2, mvmin, mvmax, bcost, bmv);
        if (bcost == ucost1 && bcost < (500 >> 4) * 4)
            return false;
        if (bcost == ucost2)
        {
            const int range = (int)(short)(merange >> 1) | 1;
            gCross(m, omv, 3, range, range, mvmin, mvmax, bcost, bmv);
            gCostMvX4(m, omv, -1, -2, 1, -2, -2, -1, 2, -1, mvmin, mvmax, bcost, bmv);
            gCostMvX4(m, omv, -2, 1, 2, 1, -1, 2, 1, 2, mvmin, mvmax, bcost, bmv);
            if (bcost == ucost2)
                return false;
            crossStart = range + 2;
        }
    }
    gCross(m, omv, crossStart, merange, merange >> 1, mvmin, mvmax, bcost, bmv);
    gCostMvX4(m, omv, -2, -2, -2, 2, 2, -2, 2, 2, mvmin, mvmax, bcost, bmv);

    /* hexagon grid */
    omv = bmv;
    int i = 1;
    do
    {
        if (4 * i > gmin(gmin(mvmax.x - omv.x, omv.x - mvmin.x), gmin(mvmax.y - omv.y, omv.y - mvmin.y)))
        {
#pragma unroll 1
            for (int j = 0; j < 16; j++)
            {
                const MV2 mv = { omv.x + g_hex4[j][0] * i, omv.y + g_hex4[j][1] * i };
                if (gInRange(mv, mvmin, mvmax))
                    gCostMv(m, mv.x, mv.y, bcost, bmv);
            }
        }
        else
        {
            /* all 16 points are measured; the range test of MIN_MV uses the UNSCALED y offset (motion.cpp:1100) */
            int best = -1;
#pragma unroll 1
            for (int k = 0; k < 16; k++)
            {
                const int mx = omv.x + g_hex4[k][0] * i, my = omv.y + g_hex4[k][1] * i;
                const int cost = m.sadFpel(mx, my) + m.mvc(mx << 2, my << 2);
                if ((omv.y + g_hex4[k][1] >= mvmin.y) & (omv.y + g_hex4[k][1] <= mvmax.y))
                    if (cost < bcost) { bcost = cost; best = k; }
            }
            if (best >= 0)
            {
                bmv.x = omv.x + i * g_hex4[best][0];
                bmv.y = omv.y + i * g_hex4[best][1];
            }
        }
    }
    while (++i <= merange >> 2);
    return gInRange(bmv, mvmin, mvmax);
}

/* ---- STAR (adapted from HM; motion.cpp:387-630 StarPatternSearch, :1157-1264) ---- */
struct StarState { int bcost; MV2 bmv; int pointNr, distance; };

/* COST_MV_PT_DIST (motion.cpp:249-261) */
template <typename Ctx>
LA_HD void gStarPt(Ctx& m, int mx, int my, int point, int dist, StarState& s)
{
    const int cost = m.sadFpel(mx, my) + m.mvc(mx << 2, my << 2);
    if (cost < s.bcost) { s.bcost = cost; s.bmv.x = mx; s.bmv.y = my; s.pointNr = point; s.distance = dist; }
}

template <typename Ctx>
LA_HD void gStarPattern(Ctx& m, MV2 mvmin, MV2 mvmax, StarState& s, int earlyExitIters, int merange)
{
    const MV2 omv = s.bmv;
    int saved = s.bcost, rounds = 0;
    {
        const int top = omv.y - 1, bottom = omv.y + 1, left = omv.x - 1, right = omv.x + 1;
        const bool inside = top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y;
        if (inside || top >= mvmin.y)    gStarPt(m, omv.x, top, 2, 1, s);
        if (inside || left >= mvmin.x)   gStarPt(m, left, omv.y, 4, 1, s);
        if (inside || right <= mvmax.x)  gStarPt(m, right, omv.y, 5, 1, s);
        if (inside || bottom <= mvmax.y) gStarPt(m, omv.x, bottom, 7, 1, s);
        if (s.bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
#pragma unroll 1
    for (int dist = 2; dist <= 8; dist <<= 1)
    {
        const int h = dist >> 1;
        const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
        const int top2 = omv.y - h, bottom2 = omv.y + h, left2 = omv.x - h, right2 = omv.x + h;
        saved = s.bcost;
        if (top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y)
        {
            gStarPt(m, omv.x, top, 2, dist, s);
            gStarPt(m, left2, top2, 1, h, s);
            gStarPt(m, right2, top2, 3, h, s);
            gStarPt(m, left, omv.y, 4, dist, s);
            gStarPt(m, right, omv.y, 5, dist, s);
            gStarPt(m, left2, bottom2, 6, h, s);
            gStarPt(m, right2, bottom2, 8, h, s);
            gStarPt(m, omv.x, bottom, 7, dist, s);
        }
        else
        {
            if (top >= mvmin.y) gStarPt(m, omv.x, top, 2, dist, s);
            if (top2 >= mvmin.y)
            {
                if (left2 >= mvmin.x) gStarPt(m, left2, top2, 1, h, s);
                if (right2 <= mvmax.x) gStarPt(m, right2, top2, 3, h, s);
            }
            if (left >= mvmin.x) gStarPt(m, left, omv.y, 4, dist, s);
            if (right <= mvmax.x) gStarPt(m, right, omv.y, 5, dist, s);
            if (bottom2 <= mvmax.y)
            {
                if (left2 >= mvmin.x) gStarPt(m, left2, bottom2, 6, h, s);
                if (right2 <= mvmax.x) gStarPt(m, right2, bottom2, 8, h, s);
            }
            if (bottom <= mvmax.y) gStarPt(m, omv.x, bottom, 7, dist, s);
        }
        if (s.bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
#pragma unroll 1
    for (int dist = 16; dist <= (int)(short)merange; dist <<= 1)
    {
        const int q = dist >> 2;
        const int top = omv.y - dist, bottom = omv.y + dist, left = omv.x - dist, right = omv.x + dist;
        saved = s.bcost;
        const bool inside = top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y;
        if (inside || top >= mvmin.y)    gStarPt(m, omv.x, top, 0, dist, s);
        if (inside || left >= mvmin.x)   gStarPt(m, left, omv.y, 0, dist, s);
        if (inside || right <= mvmax.x)  gStarPt(m, right, omv.y, 0, dist, s);
        if (inside || bottom <= mvmax.y) gStarPt(m, omv.x, bottom, 0, dist, s);
#pragma unroll 1
        for (int index = 1; index < 4; index++)
        {
            const int posYT = top + q * index, posYB = bottom - q * index, posXL = omv.x - q * index, posXR = omv.x + q * index;
            if (inside || posYT >= mvmin.y)
            {
                if (inside || posXL >= mvmin.x) gStarPt(m, posXL, posYT, 0, dist, s);
                if (inside || posXR <= mvmax.x) gStarPt(m, posXR, posYT, 0, dist, s);
            }
            if (inside || posYB <= mvmax.y)
            {
                if (inside || posXL >= mvmin.x) gStarPt(m, posXL, posYB, 0, dist, s);
                if (inside || posXR <= mvmax.x) gStarPt(m, posXR, posYB, 0, dist, s);
            }
        }
        if (s.bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
}

/* the two points next to point `pointNr` of a distance-1 result (motion.cpp:1166-1189) */
template <typename Ctx>
LA_HD void gStarTwoPoints(Ctx& m, MV2 mvmin, MV2 mvmax, StarState& s)
{
    const MV2 c = s.bmv;
    const MV2 mv1 = { c.x + g_starOffsets[(s.pointNr - 1) * 2][0], c.y + g_starOffsets[(s.pointNr - 1) * 2][1] };
    const MV2 mv2 = { c.x + g_starOffsets[(s.pointNr - 1) * 2 + 1][0], c.y + g_starOffsets[(s.pointNr - 1) * 2 + 1][1] };
    if (gInRange(mv1, mvmin, mvmax)) gCostMv(m, mv1.x, mv1.y, s.bcost, s.bmv);
    if (gInRange(mv2, mvmin, mvmax)) gCostMv(m, mv2.x, mv2.y, s.bcost, s.bmv);
}

template <typename Ctx>
LA_HD void gStar(Ctx& m, MV2 mvmin, MV2 mvmax, int merange, int& bcost, MV2& bmv)
{
    StarState s = { bcost, bmv, 0, 0 };
    gStarPattern(m, mvmin, mvmax, s, 3, merange);
    bool done = false;
    if (s.distance == 1)
    {
        if (s.pointNr)
        {
            const int saved = s.bcost;
            gStarTwoPoints(m, mvmin, mvmax, s);
            done = s.bcost == saved;
        }
        else
            done = true;
    }
    if (!done)
    {
        if (s.distance > 5)
        {
            /* raster refinement over the WHOLE vector range in steps of 5 (motion.cpp:1192-1228).  Four columns at a time
             * while they fit; the cost of the fourth is charged for the vector shifted by 3 bits instead of 2 -- the
             * reference's typo (:1219), part of its results */
#pragma unroll 1
            for (int y = mvmin.y; y <= mvmax.y; y += 5)
#pragma unroll 1
                for (int x = mvmin.x; x <= mvmax.x; x += 5)
                {
                    if (x + 15 <= mvmax.x)
                    {
#pragma unroll 1
                        for (int k = 0; k < 4; k++)
                        {
                            const int xk = x + 5 * k, sh = k == 3 ? 3 : 2;
                            const int cost = m.sadFpel(xk, y) + m.mvc(xk << sh, y << sh);
                            if (cost < s.bcost) { s.bcost = cost; s.bmv.x = xk; s.bmv.y = y; }
                        }
                        x += 15;
                    }
                    else
                        gCostMv(m, x, y, s.bcost, s.bmv);
                }
        }
        while (s.distance > 0)
        {
            s.distance = 0; s.pointNr = 0;
            gStarPattern(m, mvmin, mvmax, s, 32, merange);
            if (s.distance == 1)
            {
                if (s.pointNr) gStarTwoPoints(m, mvmin, mvmax, s);
                break;
            }
        }
    }
    bcost = s.bcost; bmv = s.bmv;
}

template <typename Ctx>
LA_HD int motionEstimateG(Ctx& m, MV2 mvmin, MV2 mvmax, MV2 qmvp, int merange, int method, MV2& out)
{
    m.mvpx = qmvp.x; m.mvpy = qmvp.y;
    const MV2 qmin = { mvmin.x << 2, mvmin.y << 2 }, qmax = { mvmax.x << 2, mvmax.y << 2 };
    MV2 pmv = { gmax(gmin(qmvp.x, qmax.x), qmin.x), gmax(gmin(qmvp.y, qmax.y), qmin.y) };
    const MV2 bestpre = pmv;
    const int bprecost = m.qpelSad(pmv.x, pmv.y);
    MV2 bmv = { (pmv.x + 2) >> 2, (pmv.y + 2) >> 2 };
    int bcost = bprecost;
    if ((pmv.x & 3) | (pmv.y & 3))
        bcost = m.sadFpel(bmv.x, bmv.y) + m.mvc(bmv.x << 2, bmv.y << 2);
    if (pmv.x | pmv.y)
    {
        const int cost = m.sadFpel(0, 0) + m.mvc(0, 0);
        if (cost < bcost)
        {
            bcost = cost;
            bmv.x = 0;
            bmv.y = gmax(gmin(0, mvmax.y), mvmin.y);
        }
    }
    pmv.x = (pmv.x + 2) >> 2; pmv.y = (pmv.y + 2) >> 2;      /* motion.cpp:839 */
    bool hexRefine = method == LA_HEX_SEARCH;
#define LA_GYOK(dy) ((bmv.y + (dy) >= mvmin.y) & (bmv.y + (dy) <= mvmax.y))
#define LA_GCOST(dx, dy) (m.sadFpel(bmv.x + (dx), bmv.y + (dy)) + m.mvc((bmv.x + (dx)) << 2, (bmv.y + (dy)) << 2))
    if (method == LA_DIA_SEARCH)
    {   /* diamond, radius 1 (motion.cpp:845-868) */
        bcost <<= 4;
        int i = merange;
        do
        {
            const int c0 = LA_GCOST(0, -1), c1 = LA_GCOST(0, 1), c2 = LA_GCOST(-1, 0), c3 = LA_GCOST(1, 0);
            if (LA_GYOK(-1)) bcost = gmin(bcost, (c0 << 4) + 1);
            if (LA_GYOK(1))  bcost = gmin(bcost, (c1 << 4) + 3);
            bcost = gmin(bcost, (c2 << 4) + 4);
            bcost = gmin(bcost, (c3 << 4) + 12);
            if (!(bcost & 15))
                break;
            bmv.x -= (int)((unsigned)bcost << 28) >> 30;
            bmv.y -= (int)((unsigned)bcost << 30) >> 30;
            bcost &= ~15;
        }
        while (--i && gInRange(bmv, mvmin, mvmax));
        bcost >>= 4;
    }
    else if (method == LA_UMH_SEARCH)
        hexRefine = gUmh(m, mvmin, mvmax, pmv, merange, bcost, bmv);
    else if (method == LA_STAR_SEARCH)
        gStar(m, mvmin, mvmax, merange, bcost, bmv);
    else if (method == LA_FULL_SEARCH)
    {   /* exhaustive (motion.cpp:1421-1466); under --hme the rectangle is the vector range cut to +-merange around ZERO */
        const int r = merange < 0 ? -merange : merange;
        const int y0 = gmax(mvmin.y, -r), y1 = gmin(mvmax.y, r), x0 = gmax(mvmin.x, -r), x1 = gmin(mvmax.x, r);
#pragma unroll 1
        for (int y = y0; y <= y1; y++)
#pragma unroll 1
            for (int x = x0; x <= x1; x++)
                gCostMv(m, x, y, bcost, bmv);
    }
    if (hexRefine)
    {   /* hexagon, radius 2 (motion.cpp:892-946), then the square refinement (:950-967) */
        int c0 = LA_GCOST(-2, 0), c1 = LA_GCOST(-1, 2), c2 = LA_GCOST(1, 2);
        bcost <<= 3;
        if (LA_GYOK(0)) bcost = gmin(bcost, (c0 << 3) + 2);
        if (LA_GYOK(2)) { bcost = gmin(bcost, (c1 << 3) + 3); bcost = gmin(bcost, (c2 << 3) + 4); }
        c0 = LA_GCOST(2, 0); c1 = LA_GCOST(1, -2); c2 = LA_GCOST(-1, -2);
        if (LA_GYOK(0)) bcost = gmin(bcost, (c0 << 3) + 5);
        if (LA_GYOK(-2)) { bcost = gmin(bcost, (c1 << 3) + 6); bcost = gmin(bcost, (c2 << 3) + 7); }
        if (bcost & 7)
        {
            int dir = (bcost & 7) - 2;
            if (LA_GYOK(g_hex2[dir + 1][1]))
            {
                bmv.x += g_hex2[dir + 1][0]; bmv.y += g_hex2[dir + 1][1];
#pragma unroll 1
                for (int i = (merange >> 1) - 1; i > 0 && gInRange(bmv, mvmin, mvmax); i--)
                {
                    c0 = LA_GCOST(g_hex2[dir + 0][0], g_hex2[dir + 0][1]);
                    c1 = LA_GCOST(g_hex2[dir + 1][0], g_hex2[dir + 1][1]);
                    c2 = LA_GCOST(g_hex2[dir + 2][0], g_hex2[dir + 2][1]);
                    bcost &= ~7;
                    if (LA_GYOK(g_hex2[dir + 0][1])) bcost = gmin(bcost, (c0 << 3) + 1);
                    if (LA_GYOK(g_hex2[dir + 1][1])) bcost = gmin(bcost, (c1 << 3) + 2);
                    if (LA_GYOK(g_hex2[dir + 2][1])) bcost = gmin(bcost, (c2 << 3) + 3);
                    if (!(bcost & 7))
                        break;
                    dir += (bcost & 7) - 2;
                    dir = g_mod6m1[dir + 1];
                    bmv.x += g_hex2[dir + 1][0]; bmv.y += g_hex2[dir + 1][1];
                }
            }
        }
        bcost >>= 3;
        int sdir = 0, c3;
        c0 = LA_GCOST(0, -1); c1 = LA_GCOST(0, 1); c2 = LA_GCOST(-1, 0); c3 = LA_GCOST(1, 0);
        if (LA_GYOK(-1)) { if (c0 < bcost) { bcost = c0; sdir = 1; } }
        if (LA_GYOK(1))  { if (c1 < bcost) { bcost = c1; sdir = 2; } }
        if (c2 < bcost) { bcost = c2; sdir = 3; }
        if (c3 < bcost) { bcost = c3; sdir = 4; }
        c0 = LA_GCOST(-1, -1); c1 = LA_GCOST(-1, 1); c2 = LA_GCOST(1, -1); c3 = LA_GCOST(1, 1);
        if (LA_GYOK(-1)) { if (c0 < bcost) { bcost = c0; sdir = 5; } }
        if (LA_GYOK(1))  { if (c1 < bcost) { bcost = c1; sdir = 6; } }
        if (LA_GYOK(-1)) { if (c2 < bcost) { bcost = c2; sdir = 7; } }
        if (LA_GYOK(1))  { if (c3 < bcost) { bcost = c3; sdir = 8; } }
        bmv.x += g_square1[sdir][0]; bmv.y += g_square1[sdir][1];
    }
#undef LA_GYOK
#undef LA_GCOST
    if (bprecost < bcost) { bmv = bestpre; bcost = bprecost; }
    else { bmv.x <<= 2; bmv.y <<= 2; }

    if (!bcost)
        bcost = m.mvc(bmv.x, bmv.y);        /* zero residual: no subpel, the cost is the vector's (motion.cpp:1490-1495) */
    else
    {   /* lowres subpel (motion.cpp:1496-1528): 4 half-pel SADs, re-measure with SATD, 4 quarter-pel SATDs */
        int bdir = 0;
#pragma unroll 1
        for (int i = 1; i <= 4; i++)
        {
            const int qx = bmv.x + g_square1[i][0] * 2, qy = bmv.y + g_square1[i][1] * 2;
            if ((qy < qmin.y) | (qy > qmax.y)) continue;
            const int cost = m.qpelSad(qx, qy) + m.mvc(qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += g_square1[bdir][0] * 2; bmv.y += g_square1[bdir][1] * 2;
        bcost = m.qpelSatd(bmv.x, bmv.y) + m.mvc(bmv.x, bmv.y);
        bdir = 0;
#pragma unroll 1
        for (int i = 1; i <= 4; i++)
        {
            const int qx = bmv.x + g_square1[i][0], qy = bmv.y + g_square1[i][1];
            if ((qy < qmin.y) | (qy > qmax.y)) continue;
            const int cost = m.qpelSatd(qx, qy) + m.mvc(qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += g_square1[bdir][0]; bmv.y += g_square1[bdir][1];
    }
    out = bmv;
    return bcost;
}

} // namespace la
