/* la_device.cuh -- device primitives shared by the lookahead kernels (sm_100a).
 *
 * Work decomposition used by every block-level kernel: ONE 8x8 LOWRES BLOCK = 8 LANES, lane r owns
 * pixel row r of the block; a warp always works on 4 blocks in lockstep (control flow is kept
 * warp-uniform with predicates), so every shuffle runs with the full mask and stays inside its
 * 8-lane group by construction (xor 1/2/4).  A row of 8 samples lives in registers as packed words
 * (2 x u32 for 8-bit, 4 x u32 for 16-bit samples) so SAD / averaging run on the packed-integer SIMD
 * path (__vsadu4/__vavgu4, __vsadu2/__vavgu2) and the Hadamard butterflies run as SWAR adds plus
 * warp shuffles.  No tensor cores: this is integer sum-of-absolute-(transformed-)differences, not a
 * contraction.
 *
 * Plane layout in HBM: the four half-pel planes are stored TILED, 8x8 samples per tile (one 128-byte
 * line for 16-bit samples, half a line for 8-bit), tiles in raster order over the padded plane.  An
 * arbitrary 8x8 block then touches at most 4 tiles (instead of 8-16 separate lines in a pitched
 * layout), and a lane fetches its row with two aligned vector loads (2 x LDG.128 / 2 x LDG.64).
 * ncu on the first (pitched, 32-bit-load) version showed the L1 tag stage at 83 % and issue at 22 %:
 * one warp-level load touched 32 different lines.  Host mirrors are de-tiled on fetch.
 *
 * Arithmetic definitions (what must be bit-exact) come from the reference's C primitives:
 *   SAD 8x8            source/common/pixel.cpp:40-56
 *   SATD 8x8 (2x 8x4)  source/common/pixel.cpp:190-261,281-297
 *   pixelavg_pp        source/common/pixel.cpp:545-557
 *   lowresMC           source/common/lowres.h:71-96
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace la {

#define LA_FULL 0xffffffffu

struct Geom
{
    int picW, picH, cW, cH;     /* full-res luma / chroma size */
    int srcPitch, srcPitchC;    /* samples per row of the staged full-res luma / chroma planes: picW / cW rounded up to 16, so
                                   that every row starts 16-byte aligned for the bulk copies of K1 */
    int w, h, bw, bh, ncu;      /* lowres plane size and 8x8 grid */
    int mx, my, stride, planeLines;
    long long planeSize, padOffset;
    int lambda, depth, nb;
    int tpr;                    /* tiles per plane row = stride / 8 */
    int rowsPerSlice;           /* cooperative search slices (x265cu_config::rows_per_slice); 0 = none */
    /* adaptive-quant block grid (calcAdaptiveQuantFrame, slicetype.cpp:452-472): blocks of aqBlock x aqBlock full-res
     * samples visited in raster order with a RUNNING index, aqW per row; the arrays hold ncuFull entries and the
     * frame means divide by ncuFull.  qg-size > 8: aqBlock 16, ncuFull = ncu.  qg-size 8: aqBlock 8, ncuFull = 4 ncu,
     * consumers address block (2 cuX + i, 2 cuY + j) as 2 cuX + i + (2 cuY + j) * 2 bw, which is NOT the running index
     * when picW / 8 is odd -- the reference's behaviour, reproduced as is */
    int aqBlock, aqW, aqH, ncuFull;
};

/* sample offset of buffer coordinate (X, Y) (margins included, X,Y >= 0) inside a tiled plane */
__host__ __device__ __forceinline__ long long tileOff(int X, int Y, int tpr)
{
    return ((long long)((Y >> 3) * tpr + (X >> 3)) << 6) + ((Y & 7) << 3) + (X & 7);
}

template <typename P> struct Row;
template <> struct Row<uint8_t>  { uint32_t v[2]; };
template <> struct Row<uint16_t> { uint32_t v[4]; };

/* 12-bit samples.  Same storage as the 10-bit ones (uint16_t) but a type of its own, so that every kernel template gets a separate
 * instantiation for them and the one primitive that differs -- the SATD, whose 4x4 Hadamard coefficients reach 16 * 4095 and no
 * longer fit the packed signed 16-bit lanes of the 8 / 10-bit version -- can be selected by type (groupSatdRows<px12> below)
 * without a run-time branch in the 8 / 10-bit kernels.  Everything else (loads, SAD, averaging, packing) is the uint16_t code. */
struct px12
{
    uint16_t v;
    px12() = default;
    __host__ __device__ __forceinline__ px12(int x) : v((uint16_t)x) {}
    __host__ __device__ __forceinline__ operator int() const { return v; }
};
using ::__ldg;
__device__ __forceinline__ px12 __ldg(const px12* p) { return px12((int)::__ldg((const uint16_t*)p)); }
template <> struct Row<px12> : Row<uint16_t>
{
    __device__ __forceinline__ Row() {}
    __device__ __forceinline__ Row(const Row<uint16_t>& b) : Row<uint16_t>(b) {}
};

/* 8 samples of row Y starting at column X of a tiled plane: two aligned vector loads + funnel shift */
__device__ __forceinline__ Row<uint8_t> loadRowT(const uint8_t* plane, int tpr, int X, int Y)
{
    const uint8_t* t = plane + (((long long)((Y >> 3) * tpr + (X >> 3))) << 6) + ((Y & 7) << 3);
    const uint2 a = __ldg((const uint2*)t);
    const uint2 b = __ldg((const uint2*)(t + 64));
    const int fx = X & 7;
    const bool s1 = fx & 4;
    const uint32_t u0 = s1 ? a.y : a.x, u1 = s1 ? b.x : a.y, u2 = s1 ? b.y : b.x;
    const uint32_t sh = (uint32_t)(fx & 3) * 8;
    Row<uint8_t> r;
    r.v[0] = __funnelshift_r(u0, u1, sh);
    r.v[1] = __funnelshift_r(u1, u2, sh);
    return r;
}

__device__ __forceinline__ Row<uint16_t> loadRowT(const uint16_t* plane, int tpr, int X, int Y)
{
    /* The 8 samples start at sample fx of a 16-byte tile row and may run into the tile to the right (+64 samples).
     * Three 8-byte loads fetch exactly the 4-sample chunks they touch (chunk i = fx >> 2 onwards), which leaves a
     * one-word select and a half-word funnel shift: 9 ALU operations instead of the 15 that picking 5 of the 8
     * words of two 16-byte loads took -- the ALU pipe is what bounds the search kernel (ncu: 65 % busy), the
     * load pipe has room.  The chunk addresses are affine in i, so they cost multiply-adds on the FMA pipe. */
    const uint16_t* t = plane + (((long long)((Y >> 3) * tpr + (X >> 3))) << 6) + ((Y & 7) << 3);
    const int fx = X & 7;
    const int i = fx >> 2;
    const uint2 c0 = __ldg((const uint2*)(t + 4 * i));             /* A0 | A1 */
    const uint2 c1 = __ldg((const uint2*)(t + 4 + 60 * i));        /* A1 | B0 */
    const uint2 c2 = __ldg((const uint2*)(t + 64 + 4 * i));        /* B0 | B1 */
    const bool s1 = fx & 2;
    const uint32_t t0 = s1 ? c0.y : c0.x, t1 = s1 ? c1.x : c0.y, t2 = s1 ? c1.y : c1.x, t3 = s1 ? c2.x : c1.y,
                   t4 = s1 ? c2.y : c2.x;
    const uint32_t sh = (uint32_t)(fx & 1) * 16;
    Row<uint16_t> r;
    r.v[0] = __funnelshift_r(t0, t1, sh);
    r.v[1] = __funnelshift_r(t1, t2, sh);
    r.v[2] = __funnelshift_r(t2, t3, sh);
    r.v[3] = __funnelshift_r(t3, t4, sh);
    return r;
}

/* row of a block whose column is a multiple of 8: one vector load */
__device__ __forceinline__ Row<uint8_t> loadRowAligned(const uint8_t* plane, int tpr, int X, int Y)
{
    const uint2 a = __ldg((const uint2*)(plane + (((long long)((Y >> 3) * tpr + (X >> 3))) << 6) + ((Y & 7) << 3)));
    Row<uint8_t> r; r.v[0] = a.x; r.v[1] = a.y;
    return r;
}
__device__ __forceinline__ Row<uint16_t> loadRowAligned(const uint16_t* plane, int tpr, int X, int Y)
{
    const uint4 a = __ldg((const uint4*)(plane + (((long long)((Y >> 3) * tpr + (X >> 3))) << 6) + ((Y & 7) << 3)));
    Row<uint16_t> r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    return r;
}

__device__ __forceinline__ Row<px12> loadRowT(const px12* plane, int tpr, int X, int Y) { return loadRowT((const uint16_t*)plane, tpr, X, Y); }
__device__ __forceinline__ Row<px12> loadRowAligned(const px12* plane, int tpr, int X, int Y) { return loadRowAligned((const uint16_t*)plane, tpr, X, Y); }

/* pitched (ordinary) memory: 8 samples at an arbitrary sample address (used by the unit-test kernel) */
__device__ __forceinline__ Row<uint8_t> loadRowPitched(const uint8_t* p)
{
    Row<uint8_t> r; r.v[0] = r.v[1] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i >> 2] |= (uint32_t)p[i] << ((i & 3) * 8);
    return r;
}
__device__ __forceinline__ Row<uint16_t> loadRowPitched(const uint16_t* p)
{
    Row<uint16_t> r; r.v[0] = r.v[1] = r.v[2] = r.v[3] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i >> 1] |= (uint32_t)p[i] << ((i & 1) * 16);
    return r;
}

__device__ __forceinline__ Row<px12> loadRowPitched(const px12* p) { return loadRowPitched((const uint16_t*)p); }

/* (a + b + 1) >> 1 per sample: pixelavg_pp */
__device__ __forceinline__ Row<uint8_t> avgRow(const Row<uint8_t>& a, const Row<uint8_t>& b)
{
    Row<uint8_t> r;
    r.v[0] = __vavgu4(a.v[0], b.v[0]);
    r.v[1] = __vavgu4(a.v[1], b.v[1]);
    return r;
}
__device__ __forceinline__ Row<uint16_t> avgRow(const Row<uint16_t>& a, const Row<uint16_t>& b)
{
    /* samples are at most 10 bits wide (x265cu_create rejects more), so the two 16-bit lanes cannot carry into
     * each other: one 3-input add, one shift, one mask */
    Row<uint16_t> r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.v[i] = ((a.v[i] + b.v[i] + 0x00010001u) >> 1) & 0x7fff7fffu;
    return r;
}

/* this lane's share (one row) of an 8x8 SAD */
__device__ __forceinline__ int sadRow(const Row<uint8_t>& a, const Row<uint8_t>& b)
{
    return (int)(__vsadu4(a.v[0], b.v[0]) + __vsadu4(a.v[1], b.v[1]));
}
/* packed 16-bit min / max are native on sm_100 (VIMNMX.U16x2); max - min never borrows across the lanes */
__device__ __forceinline__ uint32_t absdiffU16x2(uint32_t a, uint32_t b)
{
    uint32_t mx, mn;
    asm("max.u16x2 %0, %1, %2;" : "=r"(mx) : "r"(a), "r"(b));
    asm("min.u16x2 %0, %1, %2;" : "=r"(mn) : "r"(a), "r"(b));
    return mx - mn;
}
__device__ __forceinline__ int sadRow(const Row<uint16_t>& a, const Row<uint16_t>& b)
{
    /* four packed |a-b| pairs; each lane sums to at most 4 * 1023 */
    const uint32_t s = absdiffU16x2(a.v[0], b.v[0]) + absdiffU16x2(a.v[1], b.v[1]) + absdiffU16x2(a.v[2], b.v[2]) + absdiffU16x2(a.v[3], b.v[3]);
    return (int)((s & 0xffffu) + (s >> 16));
}

__device__ __forceinline__ int px(const Row<uint8_t>& r, int i)  { return (int)((r.v[i >> 2] >> ((i & 3) * 8)) & 0xffu); }
__device__ __forceinline__ int px(const Row<uint16_t>& r, int i) { return (int)((r.v[i >> 1] >> ((i & 1) * 16)) & 0xffffu); }

/* build a row from 8 sample values (compile-time indices after unrolling) */
__device__ __forceinline__ void packRow(Row<uint8_t>& r, const int v[8])
{
    r.v[0] = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
    r.v[1] = (uint32_t)v[4] | ((uint32_t)v[5] << 8) | ((uint32_t)v[6] << 16) | ((uint32_t)v[7] << 24);
}
__device__ __forceinline__ void packRow(Row<uint16_t>& r, const int v[8])
{
#pragma unroll
    for (int i = 0; i < 4; i++) r.v[i] = (uint32_t)v[2 * i] | ((uint32_t)v[2 * i + 1] << 16);
}

template <typename P>
__device__ __forceinline__ void diffRow(const Row<P>& a, const Row<P>& b, int d[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = px(a, i) - px(b, i);
}

/* sum over the 8 lanes of a group; the whole warp must call it converged */
#ifndef LA_REDUX
#define LA_REDUX 0      /* 1: REDUX.SUM over the group's 8 lanes instead of three shuffle + add stages (tuning variant) */
#endif
__device__ __forceinline__ int groupSum(int v)
{
#if LA_REDUX
    return (int)__reduce_add_sync(0xffu << (threadIdx.x & 24), (unsigned)v);
#endif
    v += __shfl_xor_sync(LA_FULL, v, 1);
    v += __shfl_xor_sync(LA_FULL, v, 2);
    v += __shfl_xor_sync(LA_FULL, v, 4);
    return v;
}

/* |lo| + (|hi| << 16) of a SWAR pair lo + (hi << 16) (the x264 abs2 trick, pixel.cpp:201-208) */
__device__ __forceinline__ uint32_t abs2(uint32_t a)
{
    const uint32_t s = ((a >> 15) & 0x10001u) * 0xffffu;
    return (a + s) ^ s;
}

/* A row re-ordered for the SWAR Hadamard: word i holds sample i in its low half and sample i+4 in its high half
 * (the x264 sum2_t pairing of pixel.cpp:248-251), so one 32-bit subtract yields two differences and one add /
 * subtract is a butterfly of both 4x4 blocks of the row at once. */
struct RowH { uint32_t v[4]; };

__device__ __forceinline__ RowH toH(const Row<uint8_t>& r)
{
    RowH h;
    h.v[0] = __byte_perm(r.v[0], r.v[1], 0x7470) & 0x00ff00ffu;     /* bytes: s0, -, s4, - */
    h.v[1] = __byte_perm(r.v[0], r.v[1], 0x7571) & 0x00ff00ffu;
    h.v[2] = __byte_perm(r.v[0], r.v[1], 0x7672) & 0x00ff00ffu;
    h.v[3] = __byte_perm(r.v[0], r.v[1], 0x7773) & 0x00ff00ffu;
    return h;
}
__device__ __forceinline__ RowH toH(const Row<uint16_t>& r)
{
    RowH h;
    h.v[0] = __byte_perm(r.v[0], r.v[2], 0x5410);      /* s0 | s4 << 16 */
    h.v[1] = __byte_perm(r.v[0], r.v[2], 0x7632);      /* s1 | s5 << 16 */
    h.v[2] = __byte_perm(r.v[1], r.v[3], 0x5410);      /* s2 | s6 << 16 */
    h.v[3] = __byte_perm(r.v[1], r.v[3], 0x7632);      /* s3 | s7 << 16 */
    return h;
}

/* 8x8 SATD of each group's block, lane r holding row r of both operands in RowH order.  The reference sums two
 * 8x4 SATDs, each = (sum |4x4 Hadamard coefficients| of its two 4x4 blocks) >> 1 (pixel.cpp:239-297).  Lanes 0-3 of a
 * group hold the upper 8x4, lanes 4-7 the lower.  Differences and horizontal butterflies are SWAR (low half = left
 * 4x4 block, high half = right one), vertical butterflies are two xor-shuffle stages, abs2 undoes the borrows.
 * Valid for |d| <= 1023 (8- and 10-bit): coefficients stay below 2^15.  Every lane returns the total. */
__device__ __forceinline__ int groupSatdH(const RowH& a, const RowH& b)
{
    const uint32_t d0 = a.v[0] - b.v[0], d1 = a.v[1] - b.v[1], d2 = a.v[2] - b.v[2], d3 = a.v[3] - b.v[3];
    const uint32_t a0 = d0 + d1, a1 = d0 - d1, a2 = d2 + d3, a3 = d2 - d3;
    uint32_t p[4] = { a0 + a2, a1 + a3, a0 - a2, a1 - a3 };
    const bool odd1 = threadIdx.x & 1, odd2 = threadIdx.x & 2;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        uint32_t t = __shfl_xor_sync(LA_FULL, p[i], 1);
        uint32_t q = odd1 ? t - p[i] : p[i] + t;
        t = __shfl_xor_sync(LA_FULL, q, 2);
        q = odd2 ? t - q : q + t;
        sum += abs2(q);
    }
    int s = (int)((sum & 0xffffu) + (sum >> 16));
    s += __shfl_xor_sync(LA_FULL, s, 1);
    s += __shfl_xor_sync(LA_FULL, s, 2);
    return (s >> 1) + (__shfl_xor_sync(LA_FULL, s, 4) >> 1);
}

template <typename P>
__device__ __forceinline__ int groupSatdRows(const Row<P>& a, const Row<P>& b)
{
    return groupSatdH(toH(a), toH(b));
}

/* The same two reductions with an explicit lane mask (the 8 lanes of ONE group): the four groups of a warp may then run
 * different control flow -- the generic searches of --hme (la_me_generic.cuh) are plain sequential code per block, not the
 * warp-wide lockstep of the default search */
__device__ __forceinline__ int groupSumM(int v, unsigned mask)
{
    v += __shfl_xor_sync(mask, v, 1);
    v += __shfl_xor_sync(mask, v, 2);
    v += __shfl_xor_sync(mask, v, 4);
    return v;
}
__device__ __forceinline__ int groupSatdHM(const RowH& a, const RowH& b, unsigned mask)
{
    const uint32_t d0 = a.v[0] - b.v[0], d1 = a.v[1] - b.v[1], d2 = a.v[2] - b.v[2], d3 = a.v[3] - b.v[3];
    const uint32_t a0 = d0 + d1, a1 = d0 - d1, a2 = d2 + d3, a3 = d2 - d3;
    uint32_t p[4] = { a0 + a2, a1 + a3, a0 - a2, a1 - a3 };
    const bool odd1 = threadIdx.x & 1, odd2 = threadIdx.x & 2;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        uint32_t t = __shfl_xor_sync(mask, p[i], 1);
        uint32_t q = odd1 ? t - p[i] : p[i] + t;
        t = __shfl_xor_sync(mask, q, 2);
        q = odd2 ? t - q : q + t;
        sum += abs2(q);
    }
    int s = (int)((sum & 0xffffu) + (sum >> 16));
    s += __shfl_xor_sync(mask, s, 1);
    s += __shfl_xor_sync(mask, s, 2);
    return (s >> 1) + (__shfl_xor_sync(mask, s, 4) >> 1);
}

/* ---- 12-bit SATD: the same two 8x4 SATDs with every butterfly in 32-bit integers (|d| <= 4095, coefficients up to 65520) ---- */
__device__ __forceinline__ void hadamardRowWide(const int d[8], int h[8])
{
#pragma unroll
    for (int k = 0; k < 8; k += 4)
    {
        const int a0 = d[k] + d[k + 1], a1 = d[k] - d[k + 1], a2 = d[k + 2] + d[k + 3], a3 = d[k + 2] - d[k + 3];
        h[k] = a0 + a2; h[k + 1] = a1 + a3; h[k + 2] = a0 - a2; h[k + 3] = a1 - a3;
    }
}
/* 8 lanes per block, lane r holds row r (groupSatdH / groupSatdHM for wide samples); `mask` names the lanes taking part */
__device__ __forceinline__ int groupSatdWide(const int d[8], unsigned mask)
{
    int h[8];
    hadamardRowWide(d, h);
    const bool odd1 = threadIdx.x & 1, odd2 = threadIdx.x & 2;
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        int t = __shfl_xor_sync(mask, h[i], 1);
        int q = odd1 ? t - h[i] : h[i] + t;
        t = __shfl_xor_sync(mask, q, 2);
        q = odd2 ? t - q : q + t;
        sum += abs(q);
    }
    int s = sum;
    s += __shfl_xor_sync(mask, s, 1);
    s += __shfl_xor_sync(mask, s, 2);
    return (s >> 1) + (__shfl_xor_sync(mask, s, 4) >> 1);
}
template <>
__device__ __forceinline__ int groupSatdRows<px12>(const Row<px12>& a, const Row<px12>& b)
{
    int d[8];
    diffRow(a, b, d);
    return groupSatdWide(d, LA_FULL);
}

/* the 8-lane SATD with an explicit lane mask, by sample type (the --hme search kernel) */
template <typename P>
__device__ __forceinline__ int groupSatdRowsM(const Row<P>& a, const Row<P>& b, unsigned mask)
{
    return groupSatdHM(toH(a), toH(b), mask);
}
template <>
__device__ __forceinline__ int groupSatdRowsM<px12>(const Row<px12>& a, const Row<px12>& b, unsigned mask)
{
    int d[8];
    diffRow(a, b, d);
    return groupSatdWide(d, mask);
}

/* ---- the same primitives for the FOUR-lanes-per-block decomposition of the motion search: lane l of a 4-lane group owns
 * rows 2l and 2l+1 of the 8x8 block, a warp works on 8 blocks.  Everything that is uniform inside a group (addresses, mv
 * costs, comparisons, the search's control flow) is then executed once per 4 lanes instead of once per 8, and the first
 * vertical Hadamard stage pairs the lane's own two rows in registers instead of going through a shuffle: ncu had the
 * 8-lane search kernel on the ALU pipe (59-65 % busy, the top pipe) with one third of its instructions group-uniform. */

/* this lane's share (two rows) of an 8x8 SAD */
__device__ __forceinline__ int sadRows2(const Row<uint8_t>& fa, const Row<uint8_t>& fb, const Row<uint8_t>& pa, const Row<uint8_t>& pb)
{
    return (int)(__vsadu4(fa.v[0], pa.v[0]) + __vsadu4(fa.v[1], pa.v[1]) + __vsadu4(fb.v[0], pb.v[0]) + __vsadu4(fb.v[1], pb.v[1]));
}
__device__ __forceinline__ int sadRows2(const Row<uint16_t>& fa, const Row<uint16_t>& fb, const Row<uint16_t>& pa, const Row<uint16_t>& pb)
{
    /* eight packed |a-b| pairs; each 16-bit lane sums to at most 8 * 1023 */
    const uint32_t s = absdiffU16x2(fa.v[0], pa.v[0]) + absdiffU16x2(fa.v[1], pa.v[1]) + absdiffU16x2(fa.v[2], pa.v[2]) + absdiffU16x2(fa.v[3], pa.v[3]) +
                       absdiffU16x2(fb.v[0], pb.v[0]) + absdiffU16x2(fb.v[1], pb.v[1]) + absdiffU16x2(fb.v[2], pb.v[2]) + absdiffU16x2(fb.v[3], pb.v[3]);
    return (int)((s & 0xffffu) + (s >> 16));
}

/* sum over the 4 lanes of a group; the whole warp must call it converged */
__device__ __forceinline__ int group4Sum(int v)
{
    v += __shfl_xor_sync(LA_FULL, v, 1);
    v += __shfl_xor_sync(LA_FULL, v, 2);
    return v;
}

/* horizontal half of the SWAR Hadamard of one row (both 4x4 blocks of the row at once) */
__device__ __forceinline__ void hadamardRowH(const RowH& a, const RowH& b, uint32_t p[4])
{
    const uint32_t d0 = a.v[0] - b.v[0], d1 = a.v[1] - b.v[1], d2 = a.v[2] - b.v[2], d3 = a.v[3] - b.v[3];
    const uint32_t a0 = d0 + d1, a1 = d0 - d1, a2 = d2 + d3, a3 = d2 - d3;
    p[0] = a0 + a2; p[1] = a1 + a3; p[2] = a0 - a2; p[3] = a1 - a3;
}

/* 8x8 SATD (two 8x4 SATDs, pixel.cpp:239-297) of each 4-lane group's block; lane l holds rows 2l (a0 / b0) and 2l+1
 * (a1 / b1) of both operands in RowH order, so lanes 0-1 hold the upper 8x4 and lanes 2-3 the lower.  Vertical
 * butterflies: rows (2l, 2l+1) in registers, then one xor-1 shuffle stage.  A lane accumulates 8 of the 16 coefficients
 * of each of its two 4x4 blocks per SWAR half: sum |c| <= 4 * sqrt(16 * 16 * 1023^2) = 65472 < 2^16 (Cauchy-Schwarz over
 * the whole 4x4 block), so one 32-bit accumulator holds both halves.  Every lane returns the total. */
__device__ __forceinline__ int group4SatdH(const RowH& a0, const RowH& b0, const RowH& a1, const RowH& b1)
{
    uint32_t p0[4], p1[4];
    hadamardRowH(a0, b0, p0);
    hadamardRowH(a1, b1, p1);
    const bool odd = threadIdx.x & 1;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const uint32_t s = p0[i] + p1[i], d = p0[i] - p1[i];
        const uint32_t ts = __shfl_xor_sync(LA_FULL, s, 1), td = __shfl_xor_sync(LA_FULL, d, 1);
        const uint32_t qs = odd ? ts - s : s + ts;
        const uint32_t qd = odd ? td - d : d + td;
        sum += abs2(qs) + abs2(qd);
    }
    int t = (int)((sum & 0xffffu) + (sum >> 16));
    t += __shfl_xor_sync(LA_FULL, t, 1);
    return (t >> 1) + (__shfl_xor_sync(LA_FULL, t, 2) >> 1);
}

template <typename P>
__device__ __forceinline__ int group4SatdRows(const Row<P>& fa, const Row<P>& fb, const Row<P>& pa, const Row<P>& pb)
{
    return group4SatdH(toH(fa), toH(pa), toH(fb), toH(pb));
}
/* 12-bit: 32-bit butterflies (see groupSatdWide); rows 2l / 2l+1 pair inside the lane, then one shuffle stage */
template <>
__device__ __forceinline__ int group4SatdRows<px12>(const Row<px12>& fa, const Row<px12>& fb, const Row<px12>& pa, const Row<px12>& pb)
{
    int d0[8], d1[8], h0[8], h1[8];
    diffRow(fa, pa, d0);
    diffRow(fb, pb, d1);
    hadamardRowWide(d0, h0);
    hadamardRowWide(d1, h1);
    const bool odd = threadIdx.x & 1;
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
        const int s = h0[i] + h1[i], d = h0[i] - h1[i];
        const int ts = __shfl_xor_sync(LA_FULL, s, 1), td = __shfl_xor_sync(LA_FULL, d, 1);
        sum += abs(odd ? ts - s : s + ts) + abs(odd ? td - d : d + td);
    }
    int t = sum;
    t += __shfl_xor_sync(LA_FULL, t, 1);
    return (t >> 1) + (__shfl_xor_sync(LA_FULL, t, 2) >> 1);
}

/* the four tiled half-pel planes of one frame and the origin of the group's block in buffer coordinates */
template <typename P>
struct RefBlock
{
    const P* plane0;        /* start of the tiled buffer of plane 0; plane i at + i * planeSize */
    long long planeSize;
    int tpr;
    int X0, Y0;             /* buffer coordinates (margins included) of the block's top-left sample */
};

/* this lane's row of the motion-compensated block: ReferencePlanes::lowresMC (lowres.h:71-96) */
template <typename P>
__device__ __forceinline__ Row<P> mcRow(const RefBlock<P>& rb, int qx, int qy, int r)
{
    const int hA = (qy & 2) | ((qx & 2) >> 1);
    const Row<P> A = loadRowT(rb.plane0 + hA * rb.planeSize, rb.tpr, rb.X0 + (qx >> 2), rb.Y0 + (qy >> 2) + r);
    if (__any_sync(LA_FULL, (qx | qy) & 1))
    {
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const Row<P> B = loadRowT(rb.plane0 + hB * rb.planeSize, rb.tpr, rb.X0 + (qx2 >> 2), rb.Y0 + (qy2 >> 2) + r);
        /* for a half/full-pel vector B == A and the rounded average returns A unchanged */
        return avgRow(A, B);
    }
    return A;
}

/* Progress counters between CTAs are polled with a RELAXED load: on this architecture ld.acquire.gpu (and
 * __threadfence) is followed by CCTL.IVALL, which throws away the SM's whole L1 on every poll (ncu on the first
 * version: 2.2e9 polls, L1 hit rate 13 %).  The data guarded by the counter is read with ld.cg, i.e. from L2, the
 * point of coherence, so no L1 invalidation is needed; the producer publishes with st.release.gpu. */
__device__ __forceinline__ int ldRelaxed(const int* p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelease(int* p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

/* x265_exp2fix8 (source/common/common.cpp:96-103, LUT constants.cpp:552-558) */
__device__ const unsigned char c_exp2_lut[64] = {
    0, 3, 6, 8, 11, 14, 17, 20, 23, 26, 29, 32, 36, 39, 42, 45,
    48, 52, 55, 58, 62, 65, 69, 72, 76, 80, 83, 87, 91, 94, 98, 102,
    106, 110, 114, 118, 122, 126, 130, 135, 139, 143, 147, 152, 156, 161, 165, 170,
    175, 179, 184, 189, 194, 198, 203, 208, 214, 219, 224, 229, 234, 240, 245, 250 };

__device__ __forceinline__ int exp2fix8(double x)
{
    /* (int)(x * (-64.f / 6.f) + 512.5f) with the float constants promoted to double, no FMA */
    const double k = (double)(-64.f / 6.f);
    int i = (int)__dadd_rn(__dmul_rn(x, k), 512.5);
    if (i < 0) return 0;
    if (i > 1023) return 0xffff;
    return (c_exp2_lut[i & 63] + 256) << (i >> 6) >> 8;
}

} // namespace la
