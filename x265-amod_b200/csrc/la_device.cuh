/* la_device.cuh -- device primitives shared by the lookahead kernels (sm_100a).
 *
 * Work decomposition used by every block-level kernel: ONE 8x8 LOWRES BLOCK = 8 LANES, lane r owns
 * pixel row r of the block.  A row of 8 samples lives in registers as packed words (2 x u32 for
 * 8-bit, 4 x u32 for 16-bit samples) so SAD / averaging run on the packed-integer SIMD path
 * (__vsadu4/__vavgu4, __vsadu2/__vavgu2) and the Hadamard butterflies run as SWAR adds plus
 * warp shuffles inside the 8-lane group.  No tensor cores: this is integer
 * sum-of-absolute-(transformed-)differences, not a contraction.
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

struct Geom
{
    int picW, picH, cW, cH;     /* full-res luma / chroma size */
    int w, h, bw, bh, ncu;      /* lowres plane size and 8x8 grid */
    int mx, my, stride, planeLines;
    long long planeSize, padOffset;
    int lambda, depth, nb;
};

template <typename P> struct Row;
template <> struct Row<uint8_t>  { uint32_t v[2]; };
template <> struct Row<uint16_t> { uint32_t v[4]; };

/* 8 samples starting at an arbitrary sample address: aligned 32-bit loads + funnel shift.
 * Reads up to 3 bytes past the row; every plane buffer is allocated with tail padding. */
__device__ __forceinline__ Row<uint8_t> loadRow(const uint8_t* p)
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
    Row<uint8_t> r;
    r.v[0] = __funnelshift_r(w0, w1, sh);
    r.v[1] = __funnelshift_r(w1, w2, sh);
    return r;
}

__device__ __forceinline__ Row<uint16_t> loadRow(const uint16_t* p)
{
    const uintptr_t a = (uintptr_t)p;
    const uint32_t* q = (const uint32_t*)(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 2) * 8;
    const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2), w3 = __ldg(q + 3), w4 = __ldg(q + 4);
    Row<uint16_t> r;
    r.v[0] = __funnelshift_r(w0, w1, sh);
    r.v[1] = __funnelshift_r(w1, w2, sh);
    r.v[2] = __funnelshift_r(w2, w3, sh);
    r.v[3] = __funnelshift_r(w3, w4, sh);
    return r;
}

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
    Row<uint16_t> r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.v[i] = __vavgu2(a.v[i], b.v[i]);
    return r;
}

/* this lane's share (one row) of an 8x8 SAD */
__device__ __forceinline__ int sadRow(const Row<uint8_t>& a, const Row<uint8_t>& b)
{
    return (int)(__vsadu4(a.v[0], b.v[0]) + __vsadu4(a.v[1], b.v[1]));
}
__device__ __forceinline__ int sadRow(const Row<uint16_t>& a, const Row<uint16_t>& b)
{
    return (int)(__vsadu2(a.v[0], b.v[0]) + __vsadu2(a.v[1], b.v[1]) + __vsadu2(a.v[2], b.v[2]) + __vsadu2(a.v[3], b.v[3]));
}

__device__ __forceinline__ int px(const Row<uint8_t>& r, int i)  { return (int)((r.v[i >> 2] >> ((i & 3) * 8)) & 0xffu); }
__device__ __forceinline__ int px(const Row<uint16_t>& r, int i) { return (int)((r.v[i >> 1] >> ((i & 1) * 16)) & 0xffffu); }

__device__ __forceinline__ void setPx(Row<uint8_t>& r, int i, int val)
{
    const int sh = (i & 3) * 8;
    r.v[i >> 2] = (r.v[i >> 2] & ~(0xffu << sh)) | ((uint32_t)val << sh);
}
__device__ __forceinline__ void setPx(Row<uint16_t>& r, int i, int val)
{
    const int sh = (i & 1) * 16;
    r.v[i >> 1] = (r.v[i >> 1] & ~(0xffffu << sh)) | ((uint32_t)val << sh);
}

template <typename P>
__device__ __forceinline__ void diffRow(const Row<P>& a, const Row<P>& b, int d[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = px(a, i) - px(b, i);
}

/* the 8 lanes of one block: lanes [8k, 8k+8) of the warp */
__device__ __forceinline__ unsigned groupMask() { return 0xFFu << ((threadIdx.x & 31u) & ~7u); }

__device__ __forceinline__ int groupSum(int v, unsigned gmask)
{
    v += __shfl_xor_sync(gmask, v, 1);
    v += __shfl_xor_sync(gmask, v, 2);
    v += __shfl_xor_sync(gmask, v, 4);
    return v;
}

/* |lo| + (|hi| << 16) of a SWAR pair lo + (hi << 16) (the x264 abs2 trick, pixel.cpp:201-208) */
__device__ __forceinline__ uint32_t abs2(uint32_t a)
{
    const uint32_t s = ((a >> 15) & 0x10001u) * 0xffffu;
    return (a + s) ^ s;
}

/* 8x8 SATD of the group's block from each lane's row of differences.  The reference sums two
 * 8x4 SATDs, each = (sum |4x4 Hadamard coefficients| of its two 4x4 blocks) >> 1.  Lanes 0-3 hold
 * the upper 8x4, lanes 4-7 the lower.  Horizontal butterflies in-lane; the two 4x4 blocks of a
 * row are packed lo/hi in one word; vertical butterflies are two xor-shuffle stages.
 * Valid for |d| <= 1023 (8- and 10-bit): coefficients stay below 2^15.  Every lane returns the total. */
__device__ __forceinline__ int groupSatd(const int d[8], unsigned gmask)
{
    const int a0 = d[0] + d[1], a1 = d[0] - d[1], a2 = d[2] + d[3], a3 = d[2] - d[3];
    const int b0 = d[4] + d[5], b1 = d[4] - d[5], b2 = d[6] + d[7], b3 = d[6] - d[7];
    uint32_t p[4];
    p[0] = (uint32_t)(a0 + a2) + ((uint32_t)(b0 + b2) << 16);
    p[1] = (uint32_t)(a1 + a3) + ((uint32_t)(b1 + b3) << 16);
    p[2] = (uint32_t)(a0 - a2) + ((uint32_t)(b0 - b2) << 16);
    p[3] = (uint32_t)(a1 - a3) + ((uint32_t)(b1 - b3) << 16);
    const bool odd1 = threadIdx.x & 1, odd2 = threadIdx.x & 2;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        uint32_t t = __shfl_xor_sync(gmask, p[i], 1);
        uint32_t q = odd1 ? t - p[i] : p[i] + t;
        t = __shfl_xor_sync(gmask, q, 2);
        q = odd2 ? t - q : q + t;
        sum += abs2(q);
    }
    int s = (int)((sum & 0xffffu) + (sum >> 16));
    s += __shfl_xor_sync(gmask, s, 1);
    s += __shfl_xor_sync(gmask, s, 2);
    return (s >> 1) + (__shfl_xor_sync(gmask, s, 4) >> 1);
}

template <typename P>
__device__ __forceinline__ int groupSatdRows(const Row<P>& a, const Row<P>& b, unsigned gmask)
{
    int d[8];
    diffRow(a, b, d);
    return groupSatd(d, gmask);
}

/* the four half-pel planes of one frame and the position of the group's block in them */
template <typename P>
struct RefBlock
{
    const P* base;          /* lowresPlane[0] + pelOffset of the block */
    long long planeSize;    /* lowresPlane[i] = lowresPlane[0] + i * planeSize */
    int stride;
};

/* this lane's row of the motion-compensated block: ReferencePlanes::lowresMC (lowres.h:71-96) */
template <typename P>
__device__ __forceinline__ Row<P> mcRow(const RefBlock<P>& rb, int qx, int qy, int r)
{
    if ((qx | qy) & 1)
    {
        const int hA = (qy & 2) | ((qx & 2) >> 1);
        const Row<P> A = loadRow(rb.base + hA * rb.planeSize + (qx >> 2) + (long long)((qy >> 2) + r) * rb.stride);
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const Row<P> B = loadRow(rb.base + hB * rb.planeSize + (qx2 >> 2) + (long long)((qy2 >> 2) + r) * rb.stride);
        return avgRow(A, B);
    }
    const int hp = (qy & 2) | ((qx & 2) >> 1);
    return loadRow(rb.base + hp * rb.planeSize + (qx >> 2) + (long long)((qy >> 2) + r) * rb.stride);
}

__device__ __forceinline__ int ldAcquire(const int* p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
