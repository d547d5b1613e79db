/* la_kernels.cuh -- the lookahead kernels (sm_100a).  See la_device.cuh for the 8-lanes-per-block
 * decomposition.  Each kernel names the reference lines whose results it reproduces.
 * Compile with -fmad=false: the few floating-point expressions must round like the reference's
 * x86-64 (no FMA) build. */
#pragma once
#include "la_device.cuh"
#include "la_me_generic.cuh"

namespace la {

#define LA_COST_MAX (1 << 28)
#define LA_LOWRES_COST_MASK 16383
#define LA_LOWRES_COST_SHIFT 14

struct FrameStatsDev            /* mirrors x265cu_frame_stats */
{
    long long costEst, costEstAq;
    unsigned long long wp_ssd[3], wp_sum[3];
    double frameVariance;
};

struct CostResultDev            /* mirrors x265cu_cost_result */
{
    long long costEst, costEstAq;
    int intraMbs, reserved;
};

/* ------------------------------------------------------------------------------------------
 * K1 (+ K2a): ONE streaming pass over the full-res picture produces the four half-pel lowres planes
 * (frame_init_lowres_core, pixel.cpp:605-628) AND the 16x16 AC energies + weightp sums of calcAdaptiveQuantFrame
 * (acEnergyCu / acEnergyPlane / pixel_var, slicetype.cpp:49-84,264-283, pixel.cpp:720-737), so the picture is read from
 * HBM once.  HBM-bound: reads 1.5 F, writes 4 P.
 *
 * Persistent CTAs, one per SM slot, walk the frame in GROUPS of LA_LR_TILES lowres tiles side by side (64 x 8 lowres
 * samples = 128 x 16 luma samples = exactly the 16x16 AQ blocks of those tiles, + their 64 x 8 U and V samples).  A group's
 * 17 luma rows x 129 columns and 2 x 8 chroma rows are staged in shared memory by BULK ASYNC COPIES (cp.async.bulk global ->
 * shared, one per row, completion counted on an mbarrier: SASS UBLKCP), issued by one thread LA_LR_STAGES groups ahead of
 * the group being processed: a ring of stages keeps several KB per CTA in flight, which is what a streaming kernel needs to
 * fill HBM, and nobody touches global memory for pixels afterwards.  Replicate clamping = PicYuv's padding (picyuv.cpp:
 * 261-285) is done once per group: vertically by pointing the copy of a row beyond the picture at the last row, horizontally
 * by a fix-up of the staged columns in the right-most groups only.  A warp owns a tile: lane (y, q) produces samples 2q, 2q+1
 * of row y of all four planes from packed shared-memory words -- FILTER(a,b,c,d) = avg(avg(a,b), avg(c,d)) with
 * avg = (x + y + 1) >> 1 is two packed average steps -- and the warp's 32 words are one contiguous tile of the tiled plane
 * layout (la_device.cuh): full-line stores.  The same lane holds a 2 x 4 piece of the tile's 16x16 source block, so the
 * block's sum / sum of squares are a warp reduction of values already in registers; lanes 0-15 add a staged chroma row each.
 * The plane margins are written by extend_border_kernel.
 * ------------------------------------------------------------------------------------------ */
template <typename P> struct Vec4;
template <> struct Vec4<uint8_t>  { typedef uchar4 T; };
template <> struct Vec4<uint16_t> { typedef ushort4 T; };
template <> struct Vec4<px12>     { typedef ushort4 T; };

#define LA_LR_TILES 8
#define LA_LR_ROW_SAMPLES 144           /* 129 needed; 144 samples = 144 / 288 bytes, a multiple of 16 for both sample sizes */
#define LA_LR_CROW_SAMPLES 64           /* chroma samples per staged row: 8 per tile */
#define LA_LR_STAGES 3
#define LA_LR_CTAS_PER_SM 5

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t avgPacked(uint32_t a, uint32_t b, uint8_t)  { return __vavgu4(a, b); }
/* 16-bit samples below 2^15: the two halves cannot carry into each other */
__device__ __forceinline__ uint32_t avgPacked(uint32_t a, uint32_t b, uint16_t) { return ((a + b + 0x00010001u) >> 1) & 0x7fff7fffu; }
__device__ __forceinline__ uint32_t avgPacked(uint32_t a, uint32_t b, px12) { return ((a + b + 0x00010001u) >> 1) & 0x7fff7fffu; }

__device__ __forceinline__ void sumSqr8s(const uint16_t* p, unsigned& sum, unsigned& sqr);
__device__ __forceinline__ void sumSqr8s(const px12* p, unsigned& sum, unsigned& sqr) { sumSqr8s((const uint16_t*)p, sum, sqr); }
/* sum / sum of squares of 8 staged samples at p (16-byte aligned for 16-bit samples, 8-byte for 8-bit) */
__device__ __forceinline__ void sumSqr8s(const uint8_t* p, unsigned& sum, unsigned& sqr)
{
    const uint2 v = *(const uint2*)p;
    sum = __vsadu4(v.x, 0) + __vsadu4(v.y, 0);
    sqr = __dp4a(v.x, v.x, __dp4a(v.y, v.y, 0u));
}
__device__ __forceinline__ void sumSqr8s(const uint16_t* p, unsigned& sum, unsigned& sqr)
{
    const uint4 v = *(const uint4*)p;
    const unsigned w[4] = { v.x, v.y, v.z, v.w };
    sum = 0; sqr = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        const unsigned lo = w[i] & 0xffffu, hi = w[i] >> 16;
        sum += lo + hi; sqr += lo * lo + hi * hi;
    }
}

template <typename P>
__global__ void __launch_bounds__(32 * LA_LR_TILES, LA_LR_CTAS_PER_SM)
lowres_fused_kernel(Geom g, const P* __restrict__ srcY, const P* __restrict__ srcU, const P* __restrict__ srcV, P* __restrict__ planes,
                    unsigned* __restrict__ energy, FrameStatsDev* stats, int doEnergy)
{
    constexpr int SPP = (int)sizeof(P);
    constexpr int ROW_BYTES = LA_LR_ROW_SAMPLES * SPP;
    constexpr int CROW_BYTES = LA_LR_CROW_SAMPLES * SPP;
    constexpr int STAGE_BYTES = 17 * ROW_BYTES + 16 * CROW_BYTES;
    __shared__ __align__(128) unsigned char s_stage[LA_LR_STAGES][STAGE_BYTES];
    __shared__ __align__(8) unsigned long long s_bar[LA_LR_STAGES];
    __shared__ unsigned long long s_acc[6];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int groupsX = (g.bw + LA_LR_TILES - 1) / LA_LR_TILES;
    const int total = groupsX * g.bh;
    const bool chroma = doEnergy && srcU != NULL;

    /* one thread arms stage `st` for group `gid` and issues its row copies */
    auto issue = [&](int gid, int st)
    {
        const int gx = gid % groupsX, ty = gid / groupsX;
        const int c0 = gx * 16 * LA_LR_TILES;
        const uint32_t bar = smemAddr(&s_bar[st]);
        const uint32_t lumaBytes = (uint32_t)(min(LA_LR_ROW_SAMPLES, g.srcPitch - c0) * SPP);
        const int cc0 = gx * LA_LR_CROW_SAMPLES;
        const uint32_t chromaBytes = chroma ? (uint32_t)(min(LA_LR_CROW_SAMPLES, g.srcPitchC - cc0) * SPP) : 0u;
        /* the stage was read (and, at the picture's right edge, patched) through the generic proxy: order that before the
         * asynchronous proxy writes into it again */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(lumaBytes * 17u + chromaBytes * 16u) : "memory");
        unsigned char* dst = s_stage[st];
#pragma unroll 1
        for (int r = 0; r < 17; r++)
        {
            const P* src = srcY + (long long)min(16 * ty + r, g.picH - 1) * g.srcPitch + c0;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smemAddr(dst + r * ROW_BYTES)), "l"(src), "r"(lumaBytes), "r"(bar) : "memory");
        }
        if (chroma)
        {
#pragma unroll 1
            for (int r = 0; r < 16; r++)
            {
                const P* src = (r < 8 ? srcU : srcV) + (long long)min(8 * ty + (r & 7), g.cH - 1) * g.srcPitchC + cc0;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smemAddr(dst + 17 * ROW_BYTES + r * CROW_BYTES)), "l"(src), "r"(chromaBytes), "r"(bar) : "memory");
            }
        }
    };

    if (tid == 0)
    {
#pragma unroll
        for (int st = 0; st < LA_LR_STAGES; st++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smemAddr(&s_bar[st])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 6) s_acc[tid] = 0;
    __syncthreads();
    if (tid == 0)
        for (int st = 0; st < LA_LR_STAGES; st++)
        {
            const int gid = blockIdx.x + st * gridDim.x;
            if (gid < total) issue(gid, st);
        }

    int it = 0;
    for (int gid = blockIdx.x; gid < total; gid += gridDim.x, it++)
    {
        const int st = it % LA_LR_STAGES;
        const uint32_t parity = (uint32_t)((it / LA_LR_STAGES) & 1);
        {
            const uint32_t bar = smemAddr(&s_bar[st]);
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        const int gx = gid % groupsX, ty = gid / groupsX;
        unsigned char* s_src = s_stage[st];
        /* columns at and beyond the picture's right edge replicate its last column (right-most groups only) */
        const int valid = g.picW - gx * 16 * LA_LR_TILES;
        const int cvalid = g.cW - gx * LA_LR_CROW_SAMPLES;
        if (valid < 16 * LA_LR_TILES + 1 || (chroma && cvalid < LA_LR_CROW_SAMPLES))
        {
            P* sp = (P*)s_src;
            if (valid < 16 * LA_LR_TILES + 1)
                for (int i = tid; i < 17 * LA_LR_ROW_SAMPLES; i += 32 * LA_LR_TILES)
                {
                    const int r = i / LA_LR_ROW_SAMPLES, x = i % LA_LR_ROW_SAMPLES;
                    if (x >= valid && x <= 16 * LA_LR_TILES) sp[r * LA_LR_ROW_SAMPLES + x] = sp[r * LA_LR_ROW_SAMPLES + valid - 1];
                }
            if (chroma && cvalid < LA_LR_CROW_SAMPLES)
            {
                P* cp = (P*)(s_src + 17 * ROW_BYTES);
                for (int i = tid; i < 16 * LA_LR_CROW_SAMPLES; i += 32 * LA_LR_TILES)
                {
                    const int r = i / LA_LR_CROW_SAMPLES, x = i % LA_LR_CROW_SAMPLES;
                    if (x >= cvalid) cp[r * LA_LR_CROW_SAMPLES + x] = cp[r * LA_LR_CROW_SAMPLES + cvalid - 1];
                }
            }
            __syncthreads();
        }
        const int tx = gx * LA_LR_TILES + warp;
        if (tx < g.bw)
        {
            const int y = lane >> 2, q = lane & 3;
            /* this lane's 3 source rows x 5 samples, starting at sample 16 * warp + 4 q of staged rows 2y .. 2y + 2 */
            const unsigned char* base = s_src + (2 * y) * ROW_BYTES + (16 * warp + 4 * q) * SPP;
            uint32_t o0, o1, o2, o3;        /* this lane's two samples of the planes (0,0), (h,0), (0,v), (h,v) */
            unsigned sum, sqr;
            if (SPP == 2)
            {
                uint32_t r[3][3];
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    const uint2 v = *(const uint2*)(base + k * ROW_BYTES);
                    r[k][0] = v.x; r[k][1] = v.y; r[k][2] = *(const unsigned short*)(base + k * ROW_BYTES + 8);
                }
                uint32_t A[2][3];
#pragma unroll
                for (int j = 0; j < 3; j++) { A[0][j] = avgPacked(r[0][j], r[1][j], (P)0); A[1][j] = avgPacked(r[1][j], r[2][j], (P)0); }
                uint32_t t[2][2];
#pragma unroll
                for (int v = 0; v < 2; v++)
                {
                    t[v][0] = avgPacked(A[v][0], __funnelshift_r(A[v][0], A[v][1], 16), (P)0);
                    t[v][1] = avgPacked(A[v][1], __funnelshift_r(A[v][1], A[v][2], 16), (P)0);
                }
                o0 = __byte_perm(t[0][0], t[0][1], 0x5410); o1 = __byte_perm(t[0][0], t[0][1], 0x7632);
                o2 = __byte_perm(t[1][0], t[1][1], 0x5410); o3 = __byte_perm(t[1][0], t[1][1], 0x7632);
                sum = 0; sqr = 0;
#pragma unroll
                for (int k = 0; k < 2; k++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
                    {
                        const unsigned lo = r[k][j] & 0xffffu, hi = r[k][j] >> 16;
                        sum += lo + hi; sqr += lo * lo + hi * hi;
                    }
            }
            else
            {
                uint32_t r[3][2];
#pragma unroll
                for (int k = 0; k < 3; k++)
                {
                    r[k][0] = *(const uint32_t*)(base + k * ROW_BYTES);
                    r[k][1] = *(const unsigned char*)(base + k * ROW_BYTES + 4);
                }
                uint32_t t[2];
#pragma unroll
                for (int v = 0; v < 2; v++)
                {
                    const uint32_t A0 = avgPacked(r[v][0], r[v + 1][0], (P)0), A1 = avgPacked(r[v][1], r[v + 1][1], (P)0);
                    t[v] = avgPacked(A0, __funnelshift_r(A0, A1, 8), (P)0);
                }
                o0 = __byte_perm(t[0], 0, 0x0020); o1 = __byte_perm(t[0], 0, 0x0031);
                o2 = __byte_perm(t[1], 0, 0x0020); o3 = __byte_perm(t[1], 0, 0x0031);
                sum = __vsadu4(r[0][0], 0) + __vsadu4(r[1][0], 0);
                sqr = __dp4a(r[0][0], r[0][0], __dp4a(r[1][0], r[1][0], 0u));
            }
            /* the warp's 32 x 2 samples are one whole tile of each plane, contiguous in the tiled layout */
            const long long tileBase = tileOff(g.mx + 8 * tx, g.my + 8 * ty, g.tpr);
            if (SPP == 2)
            {
                uint32_t* d = (uint32_t*)(planes + tileBase) + lane;
                const long long ps = g.planeSize >> 1;      /* plane stride in 32-bit words */
                d[0] = o0; d[ps] = o1; d[2 * ps] = o2; d[3 * ps] = o3;
            }
            else
            {
                unsigned short* d = (unsigned short*)(planes + tileBase) + lane;
                const long long ps = g.planeSize >> 1;      /* plane stride in 16-bit units */
                d[0] = (unsigned short)o0; d[ps] = (unsigned short)o1; d[2 * ps] = (unsigned short)o2; d[3 * ps] = (unsigned short)o3;
            }
            if (doEnergy)
            {
                /* lanes 0-7 / 8-15 add a row of the tile's co-located 8x8 U / V block */
                unsigned cs = 0, cq = 0;
                if (chroma && lane < 16)
                    sumSqr8s((const P*)(s_src + 17 * ROW_BYTES + lane * CROW_BYTES) + 8 * warp, cs, cq);
#pragma unroll
                for (int o = 16; o; o >>= 1)
                {
                    sum += __shfl_xor_sync(0xffffffffu, sum, o);
                    sqr += __shfl_xor_sync(0xffffffffu, sqr, o);
                }
#pragma unroll
                for (int o = 4; o; o >>= 1)
                {
                    cs += __shfl_xor_sync(0xffffffffu, cs, o);
                    cq += __shfl_xor_sync(0xffffffffu, cq, o);
                }
                const unsigned us = __shfl_sync(0xffffffffu, cs, 0), uq = __shfl_sync(0xffffffffu, cq, 0);
                const unsigned vs = __shfl_sync(0xffffffffu, cs, 8), vq = __shfl_sync(0xffffffffu, cq, 8);
                if (lane == 0)
                {
                    unsigned e = sqr - (unsigned)(((unsigned long long)sum * sum) >> 8);
                    atomicAdd(&s_acc[0], (unsigned long long)sum); atomicAdd(&s_acc[3], (unsigned long long)sqr);
                    if (chroma)
                    {
                        e += uq - (unsigned)(((unsigned long long)us * us) >> 6);
                        e += vq - (unsigned)(((unsigned long long)vs * vs) >> 6);
                        atomicAdd(&s_acc[1], (unsigned long long)us); atomicAdd(&s_acc[4], (unsigned long long)uq);
                        atomicAdd(&s_acc[2], (unsigned long long)vs); atomicAdd(&s_acc[5], (unsigned long long)vq);
                    }
                    energy[ty * g.bw + tx] = e;
                }
            }
        }
        __syncthreads();        /* everybody is done with this stage: refill it LA_LR_STAGES groups ahead */
        if (tid == 0)
        {
            const int next = gid + LA_LR_STAGES * gridDim.x;
            if (next < total) issue(next, st);
        }
    }
    if (doEnergy)
    {
        __syncthreads();
        if (tid < 3) atomicAdd(&stats->wp_sum[tid], s_acc[tid]);
        else if (tid < 6) atomicAdd(&stats->wp_ssd[tid - 3], s_acc[tid]);
    }
}

/* The plane margins: 4 x extendPicBorder (lowres.cpp:373-376, pixel.cpp:1044-1058) -- every margin sample replicates the
 * nearest interior sample; columns past the right margin (stride alignment) are 0 like the reference's zero-initialised
 * buffer.  One thread = one tile row (8 samples = one 16 / 8-byte store), threads in memory order. */
template <typename P>
__global__ void __launch_bounds__(256) extend_border_kernel(Geom g, P* __restrict__ planes)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long perPlane = (long long)g.tpr * g.planeLines;
    if (idx >= 4 * perPlane) return;
    const int pl = (int)(idx / perPlane);
    const long long o = idx % perPlane;
    const int tile = (int)(o >> 3), r = (int)(o & 7);
    const int X0 = (tile % g.tpr) * 8, Y = (tile / g.tpr) * 8 + r;
    if (X0 >= g.mx && X0 < g.mx + g.w && Y >= g.my && Y < g.my + g.h) return;     /* interior: written by K1 */
    P* plane = planes + pl * g.planeSize;
    typedef typename Vec4<P>::T V4;
    V4 a, b;
    P* pa = (P*)&a; P* pb = (P*)&b;
    if (X0 >= 2 * g.mx + g.w)
    {
#pragma unroll
        for (int i = 0; i < 4; i++) pa[i] = pb[i] = 0;
    }
    else
    {
        const int cy = min(max(Y, g.my), g.my + g.h - 1);
        if (X0 >= g.mx && X0 < g.mx + g.w)
        {
            const V4* s = (const V4*)(plane + tileOff(X0, cy, g.tpr));
            a = s[0]; b = s[1];
        }
        else
        {
            const P v = plane[tileOff(X0 < g.mx ? g.mx : g.mx + g.w - 1, cy, g.tpr)];
#pragma unroll
            for (int i = 0; i < 4; i++) pa[i] = pb[i] = v;
        }
    }
    V4* d = (V4*)(plane + (o << 3));
    d[0] = a; d[1] = b;
}

/* tiled -> pitched copy of the four planes, for the host mirror of Lowres::buffer[0..3] */
template <typename P>
__global__ void __launch_bounds__(256) detile_kernel(Geom g, const P* __restrict__ tiled, P* __restrict__ linear)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 4 * g.planeSize) return;
    const int pl = (int)(i / g.planeSize);
    const long long o = i % g.planeSize;
    const int Y = (int)(o / g.stride), X = (int)(o % g.stride);
    linear[i] = tiled[pl * g.planeSize + tileOff(X, Y, g.tpr)];
}

/* K2a for qg-size 8: the AC energies of 8x8 luma blocks + the co-located 4x4 chroma blocks (slicetype.cpp:64-68,79-80: var shifts
 * 6 / 4).  One 8-lane group per block: lane r sums luma row r, lanes 0-3 / 4-7 a row of the U / V block.  Blocks are
 * numbered by the reference's running index (aqW per row).  Samples are replicate-clamped like PicYuv's padding. */
template <typename P>
__global__ void __launch_bounds__(256) aq_energy8_kernel(Geom g, const P* __restrict__ y, const P* __restrict__ u,
                                                         const P* __restrict__ v, unsigned* __restrict__ energy,
                                                         FrameStatsDev* stats)
{
    __shared__ unsigned long long s_acc[6];
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0;
    __syncthreads();
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int blkRaw = blockIdx.x * 32 + grp;
    const int nblk = g.aqW * g.aqH;
    const bool act = blkRaw < nblk;
    const int blk = act ? blkRaw : nblk - 1;
    const int bx = (blk % g.aqW) * 8, by = (blk / g.aqW) * 8;
    unsigned sum = 0, sqr = 0;
    {
        const P* row = y + (long long)min(by + r, g.picH - 1) * g.srcPitch;
#pragma unroll
        for (int i = 0; i < 8; i++) { const unsigned s = row[min(bx + i, g.picW - 1)]; sum += s; sqr += s * s; }
    }
    unsigned cs = 0, cq = 0;
    if (u)
    {
        const P* row = ((r < 4) ? u : v) + (long long)min((by >> 1) + (r & 3), g.cH - 1) * g.srcPitchC;
#pragma unroll
        for (int i = 0; i < 4; i++) { const unsigned s = row[min((bx >> 1) + i, g.cW - 1)]; cs += s; cq += s * s; }
    }
    sum = groupSum((int)sum); sqr = groupSum((int)sqr);
    cs += __shfl_xor_sync(LA_FULL, cs, 1); cq += __shfl_xor_sync(LA_FULL, cq, 1);
    cs += __shfl_xor_sync(LA_FULL, cs, 2); cq += __shfl_xor_sync(LA_FULL, cq, 2);
    const unsigned us = __shfl_sync(LA_FULL, cs, 0, 8), uq = __shfl_sync(LA_FULL, cq, 0, 8);
    const unsigned vs = __shfl_sync(LA_FULL, cs, 4, 8), vq = __shfl_sync(LA_FULL, cq, 4, 8);
    if (r == 0 && act)
    {
        unsigned e = sqr - (unsigned)(((unsigned long long)sum * sum) >> 6);
        atomicAdd(&s_acc[0], (unsigned long long)sum); atomicAdd(&s_acc[3], (unsigned long long)sqr);
        if (u)
        {
            e += uq - (unsigned)(((unsigned long long)us * us) >> 4);
            e += vq - (unsigned)(((unsigned long long)vs * vs) >> 4);
            atomicAdd(&s_acc[1], (unsigned long long)us); atomicAdd(&s_acc[4], (unsigned long long)uq);
            atomicAdd(&s_acc[2], (unsigned long long)vs); atomicAdd(&s_acc[5], (unsigned long long)vq);
        }
        energy[blk] = e;
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicAdd(&stats->wp_sum[threadIdx.x], s_acc[threadIdx.x]);
    else if (threadIdx.x < 6) atomicAdd(&stats->wp_ssd[threadIdx.x - 3], s_acc[threadIdx.x]);
}

/* K2b: the transcendental finish of calcAdaptiveQuantFrame (slicetype.cpp:537-652, 681-694) for aq-mode 0..3:
 *   aq_pow_kernel    (aq-mode 2/3): qp_adj = pow(energy * bdc + 1, 0.1) per block;
 *   aq_mean_kernel   (aq-mode 2/3): the two frame sums, accumulated by ONE thread in the reference's block order
 *                    (avg_adj += qp_adj; avg_adj_pow2 += qp_adj * qp_adj, :566-567) so that every rounding of the
 *                    double sums -- and with it every qp offset and invQscaleFactor -- is the reference's by
 *                    construction.  The warp stages 1024 values at a time in shared memory (coalesced), lane 0 adds them
 *                    in order: ~10 cycles per block (two independent add chains), 0.17 ms for the 32400 blocks of a
 *                    2160p frame, on one warp, off the critical path of the search batches;
 *   aq_finish_kernel : the per-block finish, LA_AQ_CTAS CTAs.
 * n = blocks visited (aqW * aqH, the running index); the means divide by g.ncuFull (:569-570). */
/* aq-mode 4 / 5 (X265_AQ_EDGE / _BIASED): edgeFilter + computeEdge + edgeDensityCu (slicetype.cpp:98-258) in one pass, nothing
 * kept at full resolution.  One CTA = one 16x16 region of the luma plane = one AQ block (qg-size > 8) or four (qg-size 8):
 * stage the source with a 3-sample halo, 5x5 Gaussian (integer, / 159; the 2-sample picture border keeps the source) for the
 * region + 1, Sobel-like gradients of that (the 1-sample picture border keeps the SOURCE in the edge image and angle 0),
 * then the block sums.  The edge test sqrtf(gH^2 + gV^2) >= EDGE_THRESHOLD is decided in integers (the squares near the
 * threshold are far below 2^24, so the reference's float arithmetic is exact there); the angle goes through the same
 * float / double steps as the reference's (atan2 in double, rounded to float, * 180 / PI in double with its truncated PI,
 * rounded to float, truncated to a sample).  Outside the picture both images are 0 (the reference clears its buffers).
 * Output per block: the edge density (variance of the edge image, < 2^31) | mean angle within 15 degrees of a diagonal << 31.
 * Like every acEnergyVar call, the block's sum / sum of squares also goes to the weightp statistics of plane 0 (:54-55). */
template <typename P>
__global__ void __launch_bounds__(256) aq_edge_kernel(Geom g, const P* __restrict__ y, unsigned* __restrict__ edgeOut, FrameStatsDev* stats)
{
    __shared__ int s_src[22][23];
    __shared__ int s_g[18][19];
    __shared__ unsigned s_acc[4][3];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int regW = (g.picW + 15) >> 4;
    const int rx = blockIdx.x % regW, ry = blockIdx.x / regW;
    const int X0 = rx * 16, Y0 = ry * 16, W = g.picW, H = g.picH;
    if (tid < 12) s_acc[tid / 3][tid % 3] = 0;
    for (int i = tid; i < 22 * 22; i += 256)
    {
        const int r = i / 22, c = i % 22;
        const int R = Y0 - 3 + r, C = X0 - 3 + c;
        s_src[r][c] = (R >= 0 && C >= 0 && R < H && C < W) ? (int)y[(long long)R * g.srcPitch + C] : 0;
    }
    __syncthreads();
    for (int i = tid; i < 18 * 18; i += 256)
    {
        const int r = i / 18, c = i % 18;
        const int R = Y0 - 1 + r, C = X0 - 1 + c;
        int v = s_src[r + 2][c + 2];
        if (R >= 2 && C >= 2 && R < H - 2 && C < W - 2)
        {
            const int (*s)[23] = (const int (*)[23])&s_src[r][c];      /* rows r .. r + 4 = R - 2 .. R + 2 */
            const int acc = 2 * s[0][0] + 4 * s[0][1] + 5 * s[0][2] + 4 * s[0][3] + 2 * s[0][4] +
                            4 * s[1][0] + 9 * s[1][1] + 12 * s[1][2] + 9 * s[1][3] + 4 * s[1][4] +
                            5 * s[2][0] + 12 * s[2][1] + 15 * s[2][2] + 12 * s[2][3] + 5 * s[2][4] +
                            4 * s[3][0] + 9 * s[3][1] + 12 * s[3][2] + 9 * s[3][3] + 4 * s[3][4] +
                            2 * s[4][0] + 4 * s[4][1] + 5 * s[4][2] + 4 * s[4][3] + 2 * s[4][4];
            v = acc / 159;
        }
        s_g[r][c] = v;
    }
    __syncthreads();
    const int R = Y0 + ty, C = X0 + tx;
    unsigned e = 0, th = 0;
    if (R < H && C < W)
    {
        if (R >= 1 && C >= 1 && R < H - 1 && C < W - 1)
        {
            const int (*p)[19] = (const int (*)[19])&s_g[ty][tx];      /* p[1][1] is the sample itself */
            const int gH = -3 * p[0][0] + 3 * p[0][2] - 10 * p[1][0] + 10 * p[1][2] - 3 * p[2][0] + 3 * p[2][2];
            const int gV = -3 * p[0][0] - 10 * p[0][1] - 3 * p[0][2] + 3 * p[2][0] + 10 * p[2][1] + 3 * p[2][2];
            const int maxv = g.depth > 8 ? 1023 : 255;     /* EDGE_THRESHOLD (slicetype.h:64-69): 1023 for every high bit depth */
            /* |g| <= 32 * 4095: the sum of squares fits 64 bits trivially; float rounding cannot move it across maxv^2 < 2^24 */
            e = ((long long)gH * gH + (long long)gV * gV >= (long long)maxv * maxv) ? (unsigned)maxv : 0u;
            const float radians = (float)atan2((double)gV, (double)gH);
            float theta = (float)__ddiv_rn(__dmul_rn((double)radians, 180.0), 3.14159265);
            if (theta < 0) theta = __fadd_rn(180.f, theta);
            th = (unsigned)(int)theta;
        }
        else
            e = (unsigned)s_src[ty + 3][tx + 3];
    }
    const bool qg8 = g.aqBlock == 8;
    const int lane = tid & 31;
    /* a warp holds rows 2w and 2w + 1 of the region: one AQ block (16x16), or with 8x8 blocks the left half in lanes with
     * bit 3 clear and the right half in the others */
    const unsigned mask = qg8 ? ((lane & 8) ? 0xff00ff00u : 0x00ff00ffu) : 0xffffffffu;
    const unsigned sSum = __reduce_add_sync(mask, e), sSqr = __reduce_add_sync(mask, e * e), sAng = __reduce_add_sync(mask, th);
    const int q = qg8 ? ((ty >> 3) * 2 + ((tx >> 3) & 1)) : 0;
    if ((lane & (qg8 ? 23 : 31)) == 0)
    {
        atomicAdd(&s_acc[q][0], sSum); atomicAdd(&s_acc[q][1], sSqr); atomicAdd(&s_acc[q][2], sAng);
    }
    __syncthreads();
    if (tid < (qg8 ? 4 : 1))
    {
        const int bx = qg8 ? 2 * rx + (tid & 1) : rx, by = qg8 ? 2 * ry + (tid >> 1) : ry;
        if (bx < g.aqW && by < g.aqH)
        {
            const unsigned sum = s_acc[tid][0], sqr = s_acc[tid][1];
            const unsigned density = sqr - (unsigned)(((unsigned long long)sum * sum) >> (qg8 ? 6 : 8));
            const unsigned avgAngle = s_acc[tid][2] / (unsigned)(g.aqBlock * g.aqBlock);
            const bool inclined = density && ((avgAngle >= 30 && avgAngle <= 60) || (avgAngle >= 120 && avgAngle <= 150));
            edgeOut[by * g.aqW + bx] = density | (inclined ? 0x80000000u : 0u);
            atomicAdd(&stats->wp_sum[0], (unsigned long long)sum);
            atomicAdd(&stats->wp_ssd[0], (unsigned long long)sqr);
        }
    }
}

#define LA_AQ_CTAS 64
#define LA_AQ_THREADS 256

__global__ void __launch_bounds__(LA_AQ_THREADS) aq_pow_kernel(Geom g, const unsigned* __restrict__ energy,
                                                               const unsigned* __restrict__ edge /* aq-mode 4 / 5, else NULL */,
                                                               double* __restrict__ qpCuTree)
{
    const int n = g.aqW * g.aqH;
    const double bdc = (double)(1.f / (1 << (2 * (g.depth - 8))));
    for (int i = blockIdx.x * LA_AQ_THREADS + threadIdx.x; i < n; i += LA_AQ_CTAS * LA_AQ_THREADS)
    {
        unsigned e = energy[i];
        /* a block with edges takes its edge density instead of its energy (slicetype.cpp:568-585) */
        if (edge && (edge[i] & 0x7fffffffu)) e = edge[i] & 0x7fffffffu;
        qpCuTree[i] = pow(__dadd_rn(__dmul_rn((double)e, bdc), 1.0), 0.1);
    }
}

__global__ void __launch_bounds__(32) aq_mean_kernel(Geom g, const double* __restrict__ qpAdj, double* __restrict__ sums)
{
    __shared__ double s_q[1024];
    const int n = g.aqW * g.aqH, lane = threadIdx.x;
    double a = 0, b = 0;
    for (int base = 0; base < n; base += 1024)
    {
        const int m = min(1024, n - base);
        for (int i = lane; i < m; i += 32) s_q[i] = qpAdj[base + i];
        __syncwarp();
        if (lane == 0)
        {
#pragma unroll 8
            for (int i = 0; i < m; i++)
            {
                const double q = s_q[i];
                a = __dadd_rn(a, q);
                b = __dadd_rn(b, __dmul_rn(q, q));
            }
        }
        __syncwarp();
    }
    if (lane == 0) { sums[0] = a; sums[1] = b; }
}

#define LA_FADE_MAX_ROWS 1024
__global__ void __launch_bounds__(LA_AQ_THREADS) aq_finish_kernel(Geom g, const unsigned* __restrict__ energy, int aqMode,
                                                                  double aqStrength, int bWeightP, int bFades, const double* __restrict__ sums,
                                                                  double* __restrict__ qpAq, double* __restrict__ qpCuTree,
                                                                  int* __restrict__ invQ, FrameStatsDev* stats,
                                                                  const unsigned* __restrict__ edge /* aq-mode 4 / 5, else NULL */)
{
    const int tid = threadIdx.x, n = g.aqW * g.aqH;
    const bool qg8 = g.aqBlock == 8;
    const float modeOneConst = qg8 ? 11.427f : 14.427f, modeTwoConst = qg8 ? 8.f : 11.f;     /* :459-472 */
    if (blockIdx.x == 0 && bFades)
    {
        /* --fades, slicetype.cpp:697-712: frameVariance = sum over block rows of (RUNNING sum of the block variances / maxCol),
         * divided by maxRow -- the row sum is never reset and the division is an integer one, as in the reference.  With
         * weightp the loop bounds are the picture size ROUNDED to 16 (the weightp block above it reassigns maxCol / maxRow,
         * :683-684): a subset of the AQ grid (x265cu_create only accepts sizes for which it is the whole grid) */
        __shared__ unsigned long long s_row[LA_FADE_MAX_ROWS];
        const int maxCol = bWeightP ? ((g.picW + 8) >> 4) << 4 : g.picW, maxRow = bWeightP ? ((g.picH + 8) >> 4) << 4 : g.picH;
        const int nCols = (maxCol + g.aqBlock - 1) / g.aqBlock, nRows = (maxRow + g.aqBlock - 1) / g.aqBlock;
        for (int r = tid; r < nRows; r += LA_AQ_THREADS)
        {
            unsigned long long a = 0;
            for (int cc = 0; cc < nCols; cc++) a += energy[r * g.aqW + cc];
            s_row[r] = a;
        }
        __syncthreads();
        if (tid == 0)
        {
            unsigned long long run = 0;
            double fv = 0;
            for (int r = 0; r < nRows; r++) { run += s_row[r]; fv = __dadd_rn(fv, (double)(run / (unsigned long long)maxCol)); }
            stats->frameVariance = __ddiv_rn(fv, (double)maxRow);
        }
    }
    if (blockIdx.x == 0 && tid == 0 && (bWeightP || bFades))
    {
        const int maxCol = ((g.picW + 8) >> 4) << 4, maxRow = ((g.picH + 8) >> 4) << 4;
        const int wd[3] = { maxCol, maxCol >> 1, maxCol >> 1 }, ht[3] = { maxRow, maxRow >> 1, maxRow >> 1 };
        for (int i = 0; i < 3; i++)
        {
            const unsigned long long sum = stats->wp_sum[i], ssd = stats->wp_ssd[i];
            const unsigned long long area = (unsigned long long)(wd[i] * ht[i]);
            unsigned long long fsum = sum, fssd = ssd;
            if (bWeightP) fssd = ssd - (sum * sum + area / 2) / area;
            /* --fades runs acEnergyCu over every block a second time, and acEnergyVar adds each block's sum / ssd to
             * wp_sum / wp_ssd again -- after they were finalised (slicetype.cpp:54-55, 703) */
            if (bFades) { fsum += sum; fssd += ssd; }
            stats->wp_sum[i] = fsum; stats->wp_ssd[i] = fssd;
        }
    }
    if (aqMode == 0 || aqStrength == 0)
    {
        if (aqMode && aqStrength == 0)
            for (int i = blockIdx.x * LA_AQ_THREADS + tid; i < g.ncuFull; i += LA_AQ_CTAS * LA_AQ_THREADS) { qpAq[i] = 0; qpCuTree[i] = 0; invQ[i] = 256; }
        return;
    }
    double strength, avg_adj = 0, bias_strength = 0;
    if (aqMode >= 2)
    {
        const double mean = __ddiv_rn(sums[0], (double)g.ncuFull), mean2 = __ddiv_rn(sums[1], (double)g.ncuFull);
        strength = __dmul_rn(aqStrength, mean);
        avg_adj = __dadd_rn(mean, -__ddiv_rn(__dmul_rn((double)0.5f, __dadd_rn(mean2, -(double)modeTwoConst)), mean));
        bias_strength = __dmul_rn(1.0, aqStrength);
    }
    else
        strength = __dmul_rn(aqStrength, (double)1.0397f);
    for (int i = blockIdx.x * LA_AQ_THREADS + tid; i < n; i += LA_AQ_CTAS * LA_AQ_THREADS)
    {
        double qp_adj;
        if (aqMode == 3)
        {
            const double q = qpCuTree[i];
            qp_adj = __dadd_rn(__dmul_rn(strength, __dadd_rn(q, -avg_adj)),
                               __dmul_rn(bias_strength, __dadd_rn(1.0, -__ddiv_rn((double)modeTwoConst, __dmul_rn(q, q)))));
        }
        else if (aqMode == 2)
            qp_adj = __dmul_rn(strength, __dadd_rn(qpCuTree[i], -avg_adj));
        else if (aqMode == 4 || aqMode == 5)
        {
            /* slicetype.cpp:612-630: blocks whose mean gradient angle is near a diagonal and that lie above the mean get a
             * steeper slope (AQ_EDGE_BIAS 0.5); mode 5 adds a tenth of mode 3's dark bias */
            const double q = qpCuTree[i];
            const double d = __dadd_rn(q, -avg_adj);
            const bool inclined = (edge[i] >> 31) != 0;
            qp_adj = (inclined && d > 0) ? __dmul_rn(__dadd_rn(strength, 0.5), d) : __dmul_rn(strength, d);
            if (aqMode == 5)
            {
                const double dark = __ddiv_rn(__dmul_rn(bias_strength, __dadd_rn(1.0, -__ddiv_rn((double)modeTwoConst, __dmul_rn(q, q)))), (double)10.f);
                qp_adj = __dadd_rn(qp_adj, dark);
            }
        }
        else
        {
            const unsigned e = energy[i] > 1 ? energy[i] : 1;
            const float c = modeOneConst + (float)(2 * (g.depth - 8));
            qp_adj = __dmul_rn(strength, __dadd_rn(log2((double)e), -(double)c));
        }
        qpAq[i] = qp_adj;
        qpCuTree[i] = qp_adj;
        invQ[i] = exp2fix8(qp_adj);
    }
    /* entries the running index never reaches (qg-size 8 with an odd number of 8x8 block rows / columns): the reference
     * leaves them as its zero-initialised allocation had them unless cuTreeFinish of an earlier tenant of the Lowres wrote
     * them; a slot is recycled in a different order than the reference's frames, so they are put back to zero here */
    for (int i = n + blockIdx.x * LA_AQ_THREADS + tid; i < g.ncuFull; i += LA_AQ_CTAS * LA_AQ_THREADS) qpCuTree[i] = 0;
}

/* qg-size 8: the per-lowres-block scale is the mean of its four 8x8 factors (slicetype.cpp:656-670) */
__global__ void __launch_bounds__(256) aq_invq8x8_kernel(Geom g, const int* __restrict__ invQ, int* __restrict__ invQ8)
{
    const int cu = blockIdx.x * blockDim.x + threadIdx.x;
    if (cu >= g.ncu) return;
    const int cuX = cu % g.bw, cuY = cu / g.bw, fs = 2 * g.bw;
    const int i = cuX * 2 + cuY * g.bw * 4;
    invQ8[cu] = (invQ[i] + invQ[i + 1] + invQ[i + fs] + invQ[i + fs + 1]) / 4;
}

/* ------------------------------------------------------------------------------------------
 * K3: lowres intra cost.  LookaheadTLD::lowresIntraEstimate (slicetype.cpp:715-824) with
 * intraFilter<8>, intra_pred_dc_c<8>, planar_pred_c<3>, intra_pred_ang_c<8> (intrapred.cpp:31-204).
 * 8 lanes per block, 16 blocks per CTA; neighbours staged in shared memory; 12 predictions +
 * SATDs per block.  Works out of L2 (plane 0 of one frame); integer-pipe bound.
 * Control flow is warp-uniform: the data-dependent mode choices only select operands.
 * ------------------------------------------------------------------------------------------ */
__device__ const signed char c_angleTable[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
__device__ const short c_invAngleTable[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };
__device__ const unsigned char c_intraFilterFlags[35] = {
    0x38, 0x00,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38 };

/* neighbour accessor after the horizontal-mode swap of intra_pred_ang_c (intrapred.cpp:112-121) */
__device__ __forceinline__ int nbSwap(const unsigned short* s, int i, bool hor)
{
    if (!hor || i == 0) return s[i];
    return i <= 16 ? s[16 + i] : s[i - 16];
}

template <typename P>
__device__ __forceinline__ Row<P> predAngular(const unsigned short* s, int mode, int y, int maxv)
{
    const bool hor = mode < 18;
    const int angleOffset = hor ? 10 - mode : mode - 26;
    const int angle = c_angleTable[8 + angleOffset];
    const int invAngle = c_invAngleTable[angle < 0 ? -angleOffset - 1 : 0];
    int v[8];
#pragma unroll
    for (int x = 0; x < 8; x++)
    {
        const int yy = hor ? x : y, xx = hor ? y : x;       /* coordinates before the final transpose */
        int val;
        if (!angle)
        {
            val = nbSwap(s, 1 + xx, hor);
            if (xx == 0)
            {
                int t = (short)(nbSwap(s, 1, hor) + ((nbSwap(s, 17 + yy, hor) - nbSwap(s, 0, hor)) >> 1));
                val = min(max(t, 0), maxv);
            }
        }
        else
        {
            const int angleSum = (yy + 1) * angle;
            const int off = angleSum >> 5, frac = angleSum & 31;
            const int k0 = off + xx, k1 = k0 + 1;
            /* ref[k] = S(k+1) for k >= -1; projected left neighbours for k <= -2 (angle < 0) */
            const int i0 = k0 >= -1 ? k0 + 1 : 16 + ((128 + (-1 - k0) * invAngle) >> 8);
            const int i1 = k1 >= -1 ? k1 + 1 : 16 + ((128 + (-1 - k1) * invAngle) >> 8);
            const int r0 = nbSwap(s, i0, hor), r1 = nbSwap(s, i1, hor);
            val = frac ? ((32 - frac) * r0 + frac * r1 + 16) >> 5 : r0;
        }
        v[x] = val;
    }
    Row<P> out;
    packRow(out, v);
    return out;
}

template <typename P>
__global__ void __launch_bounds__(128) intra_kernel(Geom g, const P* __restrict__ plane0, const int* __restrict__ invQ,
                                                    int* __restrict__ intraCost, unsigned char* __restrict__ intraMode,
                                                    unsigned short* __restrict__ lowresCosts00, int* __restrict__ rowSatds00,
                                                    FrameStatsDev* stats)
{
    __shared__ unsigned short s_nb[16][2][34];
    __shared__ unsigned long long s_cost[2];
    if (threadIdx.x < 2) s_cost[threadIdx.x] = 0;
    __syncthreads();
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int cuRaw = blockIdx.x * 16 + grp;
    const bool act = cuRaw < g.ncu;
    const int cu = act ? cuRaw : g.ncu - 1;        /* tail groups shadow the last block and write nothing */
    const int maxv = (1 << g.depth) - 1;
    const int tpr = g.tpr;
    {
        const int cuX = cu % g.bw, cuY = cu / g.bw;
        const int X0 = g.mx + 8 * cuX, Y0 = g.my + 8 * cuY;
        const Row<P> fenc = loadRowAligned(plane0, tpr, X0, Y0 + r);
        unsigned short* nbA = s_nb[grp][0];
        unsigned short* nbF = s_nb[grp][1];
        /* top-left, 16 above, 16 left of the block (slicetype.cpp:749-752) */
        nbA[2 * r] = __ldg(plane0 + tileOff(X0 - 1 + 2 * r, Y0 - 1, tpr));
        nbA[2 * r + 1] = __ldg(plane0 + tileOff(X0 + 2 * r, Y0 - 1, tpr));
        if (r == 0) nbA[16] = __ldg(plane0 + tileOff(X0 + 15, Y0 - 1, tpr));
        nbA[17 + 2 * r] = __ldg(plane0 + tileOff(X0 - 1, Y0 + 2 * r, tpr));
        nbA[18 + 2 * r] = __ldg(plane0 + tileOff(X0 - 1, Y0 + 2 * r + 1, tpr));
        __syncwarp();
        /* intraFilter<8> (intrapred.cpp:31-51) */
        for (int i = r; i <= 32; i += 8)
        {
            int f;
            if (i == 0) f = ((nbA[0] << 1) + nbA[1] + nbA[17] + 2) >> 2;
            else if (i == 16 || i == 32) f = nbA[i];
            else if (i == 17) f = ((nbA[17] << 1) + nbA[0] + nbA[18] + 2) >> 2;
            else f = ((nbA[i] << 1) + nbA[i - 1] + nbA[i + 1] + 2) >> 2;
            nbF[i] = (unsigned short)f;
        }
        __syncwarp();

        int cost, icost = LA_COST_MAX, ilow = 0;
        {   /* DC with edge filter (intrapred.cpp:53-85) */
            int dc = groupSum(nbA[1 + r] + nbA[17 + r]) + 8;
            dc = dc / 16;
            int v[8];
#pragma unroll
            for (int x = 0; x < 8; x++)
            {
                int t = dc;
                if (r == 0) t = x == 0 ? (nbA[1] + nbA[17] + 2 * dc + 2) >> 2 : (nbA[1 + x] + 3 * dc + 2) >> 2;
                else if (x == 0) t = (nbA[17 + r] + 3 * dc + 2) >> 2;
                v[x] = t;
            }
            Row<P> pred; packRow(pred, v);
            cost = groupSatdRows(fenc, pred);
            if (cost < icost) { icost = cost; ilow = 1; }
        }
        {   /* planar on the filtered neighbours (intrapred.cpp:87-100) */
            const int topRight = nbF[9], bottomLeft = nbF[25], left = nbF[17 + r];
            int v[8];
#pragma unroll
            for (int x = 0; x < 8; x++)
                v[x] = ((7 - x) * left + (7 - r) * nbF[1 + x] + (x + 1) * topRight + (r + 1) * bottomLeft + 8) >> 4;
            Row<P> pred; packRow(pred, v);
            cost = groupSatdRows(fenc, pred);
            if (cost < icost) { icost = cost; ilow = 0; }
        }
        int acost = LA_COST_MAX, alow = 4;
        for (int mode = 5; mode < 35; mode += 5)
        {
            const Row<P> pred = predAngular<P>((c_intraFilterFlags[mode] & 8) ? nbF : nbA, mode, r, maxv);
            cost = groupSatdRows(fenc, pred);
            if (cost < acost) { acost = cost; alow = mode; }
        }
        for (int dist = 2; dist >= 1; dist--)
        {
            const int minusmode = alow - dist, plusmode = alow + dist;
            Row<P> pred = predAngular<P>((c_intraFilterFlags[minusmode] & 8) ? nbF : nbA, minusmode, r, maxv);
            cost = groupSatdRows(fenc, pred);
            if (cost < acost) { acost = cost; alow = minusmode; }
            pred = predAngular<P>((c_intraFilterFlags[plusmode] & 8) ? nbF : nbA, plusmode, r, maxv);
            cost = groupSatdRows(fenc, pred);
            if (cost < acost) { acost = cost; alow = plusmode; }
        }
        if (acost < icost) { icost = acost; ilow = alow; }
        icost += 5 * g.lambda + 4;
        if (r == 0 && act)
        {
            lowresCosts00[cu] = (unsigned short)min(icost, LA_LOWRES_COST_MASK);
            intraCost[cu] = icost;
            intraMode[cu] = (unsigned char)ilow;
            const bool scored = (cuX > 0 && cuX < g.bw - 1 && cuY > 0 && cuY < g.bh - 1) || g.bw <= 2 || g.bh <= 2;
            const int icostAq = (scored && invQ) ? ((icost * invQ[cu] + 128) >> 8) : icost;
            if (scored)
            {
                atomicAdd(&s_cost[0], (unsigned long long)icost);
                atomicAdd(&s_cost[1], (unsigned long long)icostAq);
            }
            atomicAdd(&rowSatds00[cuY], icostAq);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        atomicAdd((unsigned long long*)&stats->costEst, s_cost[0]);
        atomicAdd((unsigned long long*)&stats->costEstAq, s_cost[1]);
    }
}

/* ------------------------------------------------------------------------------------------
 * K4: motion search of one (frame, list, distance) over the whole frame.
 * Search half of CostEstimateGroup::estimateCUCost (slicetype.cpp:4103-4183) and
 * MotionEstimate::motionEstimate (motion.cpp:764-821 start, 870-969 HEX + square refine,
 * 1473-1528 lowres subpel), merange 16, subpelRefine 1.
 *
 * Dependencies: block (x,y) takes MV predictors from (x+1,y), (x,y+1), (x-1,y+1), (x+1,y+1)
 * (reverse raster order), so a frame is a wavefront.  A job is cut into STRIPS of LA_STRIP_ROWS = 4 rows and one
 * WARP (= one 32-thread CTA) owns a strip: its 4 groups search 4 consecutive rows in lockstep, row j running 2
 * columns behind row j-1.  Inside the warp the predictors travel through registers: every group keeps its last
 * three results and the group above reads them with one shuffle each (no shared memory, no barrier).  Strips of
 * a job are chained through a progress counter in global memory (st.release / relaxed polling, MVs re-read from L2
 * with ld.cg); a strip is ~250 steps of ~10-20 us, so one L2 round trip per step is noise.  Warps take their
 * (strip, job) from a ticket counter, strip-major, so a strip never waits on a warp that has not started.
 * The first version used 4-warp CTAs with one __syncthreads per wavefront step: ncu showed 45 % of all warp
 * cycles stalled on that barrier (every step cost the slowest of 16 block searches); independent warps only
 * ever wait for the strip below.  Throughput comes from running many independent jobs concurrently.
 *
 * The search is data dependent (MVP choice, hexagon iterations, early exits); it is written with per-group
 * predicates and warp-votes so the warp never diverges, which keeps every shuffle on the full mask.
 * ------------------------------------------------------------------------------------------ */
#define LA_STRIP_ROWS 4             /* 8-lane decomposition; the 4-lane one has 8 */
#ifndef LA_SEARCH4_MIN_CTAS
#define LA_SEARCH4_MIN_CTAS 20      /* 4-lane decomposition: two rows per lane want ~96 registers */
#endif
#ifndef LA_SEARCH_MIN_CTAS
#define LA_SEARCH_MIN_CTAS 28       /* resident one-warp CTAs per SM the register allocation must allow: 72 registers, no spills;
                                       32 (64 registers) spills and measured 3 % slower; 28 also leaves block slots for the short
                                       high-priority kernels */
#endif

template <typename P>
struct SearchJobDev
{
    const P* fenc0;         /* tiled buffer of lowresPlane[0] of the frame being searched */
    const P* ref0;          /* tiled buffer of plane 0 of the (possibly weighted) reference; planes at +i*planeSize */
    int*     mvOut;         /* ncu packed MVs: (x & 0xffff) | (y << 16), quarter-pel */
    int*     costOut;       /* ncu */
    int*     flagOut;       /* set to 1 when any block took the zero-MV skip (slicetype.cpp:4177-4181) */
    const int* cond;        /* NULL, or the flagOut of another search: run only if that search skipped somewhere */
    int      bidir;
    int      sliced;        /* search as cooperative slices of g.rowsPerSlice rows */
};

/* (struct MV2 lives in la_me_generic.cuh) */

/* The candidate evaluators of the search.  Inlined at each of their ~25 call sites the kernel is 75 KB of SASS, and ncu
 * at full occupancy shows "no instruction" as the largest warp stall (4.2 of 13 stalled warps per issue slot).  Built as
 * real functions (-DLA_ME_CALLS=1) the kernel shrinks to 28 KB, but the calls cost more than the instruction cache
 * gives back: measured 8 % more search time on B200 (109 vs 101 ms per 120-frame 2160p step), so inlining stays. */
#ifndef LA_ME_CALLS
#define LA_ME_CALLS 0
#endif
#if LA_ME_CALLS
#define LA_ME_FN __noinline__
#else
#define LA_ME_FN __forceinline__
#endif

/* full-pel SAD of the group's block against the reference block at buffer position (X, Y0 + r) */
template <typename P>
__device__ LA_ME_FN int sadFpelFn(Row<P> fenc, const P* plane0, int tpr, int X, int Yr)
{
    return groupSum(sadRow(fenc, loadRowT(plane0, tpr, X, Yr)));
}

/* three of them at once (independent loads in flight together); offsets packed as (d + 8) nibbles x0 y0 x1 y1 x2 y2 */
#define LA_PK3(x0, y0, x1, y1, x2, y2) \
    (((x0) + 8) | (((y0) + 8) << 4) | (((x1) + 8) << 8) | (((y1) + 8) << 12) | (((x2) + 8) << 16) | (((y2) + 8) << 20))
#ifndef LA_SAD3_LOOP
#define LA_SAD3_LOOP 0
#endif
#ifndef LA_MVP_LOOP
#define LA_MVP_LOOP 1       /* the four predictor SATDs as a real loop: 7 KB less SASS, measured 5-8 % less search time */
#endif
#ifndef LA_ME_COMPACT
#define LA_ME_COMPACT 0     /* 1: searchBlockCompact -- one loop, two evaluation sites.  30 KB of SASS instead of 78 KB and no
                               "no instruction" stalls left, but 50 % more executed instructions (39 M against 26 M per job):
                               measured 25 % SLOWER on B200 (68 against 54 us per job).  Kept as a parity-tested variant. */
#endif
template <typename P>
__device__ LA_ME_FN int3 sad3FpelFn(Row<P> fenc, const P* plane0, int tpr, int X, int Yr, int pk)
{
#if LA_SAD3_LOOP
    /* one copy of the row fetch + SAD in the instruction stream instead of three (the kernel's SASS does not fit the
     * instruction cache: "no instruction" is its largest stall) */
    int p[3];
#pragma unroll 1
    for (int i = 0; i < 3; i++, pk >>= 8)
        p[i] = groupSum(sadRow(fenc, loadRowT(plane0, tpr, X + ((pk & 15) - 8), Yr + (((pk >> 4) & 15) - 8))));
    return make_int3(p[0], p[1], p[2]);
#else
    const int p0 = sadRow(fenc, loadRowT(plane0, tpr, X + ((pk & 15) - 8), Yr + (((pk >> 4) & 15) - 8)));
    const int p1 = sadRow(fenc, loadRowT(plane0, tpr, X + (((pk >> 8) & 15) - 8), Yr + (((pk >> 12) & 15) - 8)));
    const int p2 = sadRow(fenc, loadRowT(plane0, tpr, X + (((pk >> 16) & 15) - 8), Yr + (((pk >> 20) & 15) - 8)));
    return make_int3(groupSum(p0), groupSum(p1), groupSum(p2));
#endif
}

/* lowresQPelCost (lowres.h:98-124): SAD or SATD of the motion-compensated block at quarter-pel (qx, qy) */
template <typename P>
__device__ LA_ME_FN int qpelCostFn(Row<P> fenc, const P* plane0, long long planeSize, int tpr, int X0, int Y0, int r, int qx, int qy,
                                   bool satd)
{
    const RefBlock<P> rb = { plane0, planeSize, tpr, X0, Y0 };
    const Row<P> p = mcRow(rb, qx, qy, r);
    if (satd) return groupSatdRows(fenc, p);      /* uniform */
    return groupSum(sadRow(fenc, p));
}

template <typename P>
struct MeCtx
{
    Row<P> fenc;
    RefBlock<P> rb;
    const unsigned short* mvcost;   /* centre */
    int mvpx, mvpy;
    int r;

    __device__ __forceinline__ void init(int lane) { r = lane & 7; }
    __device__ __forceinline__ void loadFenc(const P* fenc0) { fenc = loadRowAligned(fenc0, rb.tpr, rb.X0, rb.Y0 + r); }
    __device__ __forceinline__ int mvc(int qx, int qy) const
    {
        return (int)(unsigned short)(__ldg(mvcost + (qx - mvpx)) + __ldg(mvcost + (qy - mvpy)));
    }
    __device__ __forceinline__ int sadFpel(int x, int y) const { return sadFpelFn<P>(fenc, rb.plane0, rb.tpr, rb.X0 + x, rb.Y0 + y + r); }
    __device__ __forceinline__ int3 sad3Fpel(int x, int y, int pk) const { return sad3FpelFn<P>(fenc, rb.plane0, rb.tpr, rb.X0 + x, rb.Y0 + y + r, pk); }
    __device__ __forceinline__ int qpelSad(int qx, int qy) const { return qpelCostFn<P>(fenc, rb.plane0, rb.planeSize, rb.tpr, rb.X0, rb.Y0, r, qx, qy, false); }
    __device__ __forceinline__ int qpelSatd(int qx, int qy) const { return qpelCostFn<P>(fenc, rb.plane0, rb.planeSize, rb.tpr, rb.X0, rb.Y0, r, qx, qy, true); }
    __device__ __forceinline__ int qpelCost(int qx, int qy, bool satd) const { return qpelCostFn<P>(fenc, rb.plane0, rb.planeSize, rb.tpr, rb.X0, rb.Y0, r, qx, qy, satd); }
};

/* ---- the same evaluators for the 4-lanes-per-block decomposition (la_device.cuh): the lane owns rows r and r + 1 ---- */
template <typename P>
__device__ __forceinline__ int sadFpelFn4(Row<P> fa, Row<P> fb, const P* plane0, int tpr, int X, int Yr)
{
    return group4Sum(sadRows2(fa, fb, loadRowT(plane0, tpr, X, Yr), loadRowT(plane0, tpr, X, Yr + 1)));
}

template <typename P>
__device__ __forceinline__ int3 sad3FpelFn4(Row<P> fa, Row<P> fb, const P* plane0, int tpr, int X, int Yr, int pk)
{
    const int x0 = X + ((pk & 15) - 8), y0 = Yr + (((pk >> 4) & 15) - 8);
    const int x1 = X + (((pk >> 8) & 15) - 8), y1 = Yr + (((pk >> 12) & 15) - 8);
    const int x2 = X + (((pk >> 16) & 15) - 8), y2 = Yr + (((pk >> 20) & 15) - 8);
    const int p0 = sadRows2(fa, fb, loadRowT(plane0, tpr, x0, y0), loadRowT(plane0, tpr, x0, y0 + 1));
    const int p1 = sadRows2(fa, fb, loadRowT(plane0, tpr, x1, y1), loadRowT(plane0, tpr, x1, y1 + 1));
    const int p2 = sadRows2(fa, fb, loadRowT(plane0, tpr, x2, y2), loadRowT(plane0, tpr, x2, y2 + 1));
    /* a whole 8x8 SAD is at most 64 * 1023 < 2^16: two of the three sums share one register through the reduction */
    const int w = group4Sum(p0 | (p1 << 16));
    return make_int3(w & 0xffff, (int)((unsigned)w >> 16), group4Sum(p2));
}

template <typename P>
__device__ __forceinline__ int qpelCostFn4(Row<P> fa, Row<P> fb, const P* plane0, long long planeSize, int tpr, int X0, int Y0, int r2,
                                          int qx, int qy, bool satd)
{
    /* ReferencePlanes::lowresMC (lowres.h:71-96), two rows */
    const int hA = (qy & 2) | ((qx & 2) >> 1);
    const P* pA = plane0 + hA * planeSize;
    const int xa = X0 + (qx >> 2), ya = Y0 + (qy >> 2) + r2;
    Row<P> a0 = loadRowT(pA, tpr, xa, ya), a1 = loadRowT(pA, tpr, xa, ya + 1);
    if (__any_sync(LA_FULL, (qx | qy) & 1))
    {
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const P* pB = plane0 + hB * planeSize;
        const int xb = X0 + (qx2 >> 2), yb = Y0 + (qy2 >> 2) + r2;
        /* for a half/full-pel vector B == A and the rounded average returns A unchanged */
        a0 = avgRow(a0, loadRowT(pB, tpr, xb, yb));
        a1 = avgRow(a1, loadRowT(pB, tpr, xb, yb + 1));
    }
    if (satd) return group4SatdRows(fa, fb, a0, a1);      /* uniform */
    return group4Sum(sadRows2(fa, fb, a0, a1));
}

template <typename P>
struct MeCtx4
{
    Row<P> fa, fb;                  /* rows r and r + 1 of the source block */
    RefBlock<P> rb;
    const unsigned short* mvcost;   /* centre */
    int mvpx, mvpy;
    int r;                          /* first of this lane's two rows: 0, 2, 4, 6 */

    __device__ __forceinline__ void init(int lane) { r = (lane & 3) * 2; }
    __device__ __forceinline__ void loadFenc(const P* fenc0)
    {
        fa = loadRowAligned(fenc0, rb.tpr, rb.X0, rb.Y0 + r);
        fb = loadRowAligned(fenc0, rb.tpr, rb.X0, rb.Y0 + r + 1);
    }
    __device__ __forceinline__ int mvc(int qx, int qy) const
    {
        return (int)(unsigned short)(__ldg(mvcost + (qx - mvpx)) + __ldg(mvcost + (qy - mvpy)));
    }
    __device__ __forceinline__ int sadFpel(int x, int y) const { return sadFpelFn4<P>(fa, fb, rb.plane0, rb.tpr, rb.X0 + x, rb.Y0 + y + r); }
    __device__ __forceinline__ int3 sad3Fpel(int x, int y, int pk) const { return sad3FpelFn4<P>(fa, fb, rb.plane0, rb.tpr, rb.X0 + x, rb.Y0 + y + r, pk); }
    __device__ __forceinline__ int qpelSad(int qx, int qy) const { return qpelCostFn4<P>(fa, fb, rb.plane0, rb.planeSize, rb.tpr, rb.X0, rb.Y0, r, qx, qy, false); }
    __device__ __forceinline__ int qpelSatd(int qx, int qy) const { return qpelCostFn4<P>(fa, fb, rb.plane0, rb.planeSize, rb.tpr, rb.X0, rb.Y0, r, qx, qy, true); }
    __device__ __forceinline__ int qpelCost(int qx, int qy, bool satd) const { return qpelCostFn4<P>(fa, fb, rb.plane0, rb.planeSize, rb.tpr, rb.X0, rb.Y0, r, qx, qy, satd); }
};

template <typename P, int LPB> struct MeSel;
template <typename P> struct MeSel<P, 8> { typedef MeCtx<P> T; };
template <typename P> struct MeSel<P, 4> { typedef MeCtx4<P> T; };

__device__ const signed char c_hex2[8][2] = { {-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0} };
__device__ const unsigned char c_mod6m1[8] = { 5, 0, 1, 2, 3, 4, 5, 0 };
__device__ const signed char c_square1[9][2] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1} };

/* The whole warp calls this converged; every value is uniform inside an 8-lane group. */
template <typename Ctx>
__device__ __forceinline__ int motionEstimate(Ctx& m, MV2 mvmin, MV2 mvmax, MV2 qmvp, MV2& out)
{
    const int merange = 16;
    m.mvpx = qmvp.x; m.mvpy = qmvp.y;
    const MV2 qmin = { mvmin.x << 2, mvmin.y << 2 }, qmax = { mvmax.x << 2, mvmax.y << 2 };
    const MV2 pmv = { max(min(qmvp.x, qmax.x), qmin.x), max(min(qmvp.y, qmax.y), qmin.y) };
    const MV2 bestpre = pmv;
    const int bprecost = m.qpelSad(pmv.x, pmv.y);
    MV2 bmv = { (pmv.x + 2) >> 2, (pmv.y + 2) >> 2 };
    int bcost = bprecost;
    const bool sub = ((pmv.x & 3) | (pmv.y & 3)) != 0;
    if (__any_sync(LA_FULL, sub))
    {
        const int c = m.sadFpel(bmv.x, bmv.y) + m.mvc(bmv.x << 2, bmv.y << 2);
        if (sub) bcost = c;
    }
    const bool nz = (pmv.x | pmv.y) != 0;
    if (__any_sync(LA_FULL, nz))
    {
        const int cost = m.sadFpel(0, 0) + m.mvc(0, 0);
        if (nz && cost < bcost)
        {
            bcost = cost;
            bmv.x = 0;
            bmv.y = max(min(0, mvmax.y), mvmin.y);
        }
    }
#define LA_YOK(dy) ((bmv.y + (dy) >= mvmin.y) & (bmv.y + (dy) <= mvmax.y))
#define LA_COST3(c0, c1, c2, x0, y0, x1, y1, x2, y2) \
    { \
        const int3 p = m.sad3Fpel(bmv.x, bmv.y, LA_PK3(x0, y0, x1, y1, x2, y2)); \
        c0 = p.x + m.mvc((bmv.x + (x0)) << 2, (bmv.y + (y0)) << 2); \
        c1 = p.y + m.mvc((bmv.x + (x1)) << 2, (bmv.y + (y1)) << 2); \
        c2 = p.z + m.mvc((bmv.x + (x2)) << 2, (bmv.y + (y2)) << 2); \
    }
    {   /* hexagon, motion.cpp:892-946 */
        int c0, c1, c2;
        LA_COST3(c0, c1, c2, -2, 0, -1, 2, 1, 2);
        bcost <<= 3;
        if (LA_YOK(0)) { bcost = min(bcost, (c0 << 3) + 2); }
        if (LA_YOK(2)) { bcost = min(bcost, (c1 << 3) + 3); bcost = min(bcost, (c2 << 3) + 4); }
        LA_COST3(c0, c1, c2, 2, 0, 1, -2, -1, -2);
        if (LA_YOK(0)) { bcost = min(bcost, (c0 << 3) + 5); }
        if (LA_YOK(-2)) { bcost = min(bcost, (c1 << 3) + 6); bcost = min(bcost, (c2 << 3) + 7); }
        int dir = 0, iter = (merange >> 1) - 1;
        bool go = (bcost & 7) != 0;
        if (go)
        {
            dir = (bcost & 7) - 2;
            go = LA_YOK(c_hex2[dir + 1][1]);
            if (go) { bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1]; }
        }
        go = go && iter > 0 && bmv.x >= mvmin.x && bmv.x <= mvmax.x && bmv.y >= mvmin.y && bmv.y <= mvmax.y;
        while (__any_sync(LA_FULL, go))
        {
            /* half hexagon around bmv; groups that already stopped re-measure harmlessly and ignore the result */
            const int x0 = c_hex2[dir][0], y0 = c_hex2[dir][1], x1 = c_hex2[dir + 1][0], y1 = c_hex2[dir + 1][1],
                      x2 = c_hex2[dir + 2][0], y2 = c_hex2[dir + 2][1];
            LA_COST3(c0, c1, c2, x0, y0, x1, y1, x2, y2);
            if (go)
            {
                bcost &= ~7;
                if (LA_YOK(y0)) bcost = min(bcost, (c0 << 3) + 1);
                if (LA_YOK(y1)) bcost = min(bcost, (c1 << 3) + 2);
                if (LA_YOK(y2)) bcost = min(bcost, (c2 << 3) + 3);
                if (!(bcost & 7))
                    go = false;
                else
                {
                    dir += (bcost & 7) - 2;
                    dir = c_mod6m1[dir + 1];
                    bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1];
                    iter--;
                    go = iter > 0 && bmv.x >= mvmin.x && bmv.x <= mvmax.x && bmv.y >= mvmin.y && bmv.y <= mvmax.y;
                }
            }
        }
        bcost >>= 3;
    }
    {   /* square refine, motion.cpp:950-967 */
        int dir = 0, c0, c1, c2, c3;
        LA_COST3(c0, c1, c2, 0, -1, 0, 1, -1, 0);
        c3 = m.sadFpel(bmv.x + 1, bmv.y) + m.mvc((bmv.x + 1) << 2, bmv.y << 2);
        if (LA_YOK(-1)) { if (c0 < bcost) { bcost = c0; dir = 1; } }
        if (LA_YOK(1))  { if (c1 < bcost) { bcost = c1; dir = 2; } }
        if (c2 < bcost) { bcost = c2; dir = 3; }
        if (c3 < bcost) { bcost = c3; dir = 4; }
        LA_COST3(c0, c1, c2, -1, -1, -1, 1, 1, -1);
        c3 = m.sadFpel(bmv.x + 1, bmv.y + 1) + m.mvc((bmv.x + 1) << 2, (bmv.y + 1) << 2);
        if (LA_YOK(-1)) { if (c0 < bcost) { bcost = c0; dir = 5; } }
        if (LA_YOK(1))  { if (c1 < bcost) { bcost = c1; dir = 6; } }
        if (LA_YOK(-1)) { if (c2 < bcost) { bcost = c2; dir = 7; } }
        if (LA_YOK(1))  { if (c3 < bcost) { bcost = c3; dir = 8; } }
        bmv.x += c_square1[dir][0]; bmv.y += c_square1[dir][1];
    }
#undef LA_YOK
#undef LA_COST3
    if (bprecost < bcost) { bmv = bestpre; bcost = bprecost; }
    else { bmv.x <<= 2; bmv.y <<= 2; }

    /* zero residual at the start point: no subpel, cost = mvcost (motion.cpp:1490-1495) */
    const bool doSub = bcost != 0;
    const int zeroCost = m.mvc(bmv.x, bmv.y);
    if (__any_sync(LA_FULL, doSub))
    {   /* lowres subpel, motion.cpp:1496-1528: 4 half-pel SADs, re-measure with SATD, 4 quarter-pel SATDs */
        int bdir = 0;
#pragma unroll 1
        for (int i = 1; i <= 4; i++)
        {
            const int qx = bmv.x + c_square1[i][0] * 2, qy = bmv.y + c_square1[i][1] * 2;
            const bool ok = doSub && !((qy < qmin.y) | (qy > qmax.y));
            const int cost = m.qpelSad(qx, qy) + m.mvc(qx, qy);
            if (ok && cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += c_square1[bdir][0] * 2; bmv.y += c_square1[bdir][1] * 2;
        const int c = m.qpelSatd(bmv.x, bmv.y) + m.mvc(bmv.x, bmv.y);
        if (doSub) bcost = c;
        bdir = 0;
#pragma unroll 1
        for (int i = 1; i <= 4; i++)
        {
            const int qx = bmv.x + c_square1[i][0], qy = bmv.y + c_square1[i][1];
            const bool ok = doSub && !((qy < qmin.y) | (qy > qmax.y));
            const int cost = m.qpelSatd(qx, qy) + m.mvc(qx, qy);
            if (ok && cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += c_square1[bdir][0]; bmv.y += c_square1[bdir][1];
    }
    if (!doSub) bcost = zeroCost;
    out = bmv;
    return bcost;
}

__device__ __forceinline__ MV2 unpackMv(int p) { MV2 m = { (int)(short)(p & 0xffff), p >> 16 }; return m; }
__device__ __forceinline__ int packMv(MV2 m) { return (m.x & 0xffff) | (m.y << 16); }

/* ------------------------------------------------------------------------------------------
 * The whole per-block search (predictor choice + motionEstimate) as ONE loop with TWO candidate-evaluation sites.
 *
 * ncu on the inlined version above (63 KB of hot SASS: ~25 inlined candidate evaluators, each executed once per wavefront
 * step) shows "no instruction" as the largest warp stall: the SM's instruction caches are 6 KB (L0) and 32 KB (L1.5)
 * (B300_MICROARCH.md), the 28 resident warps sit at 28 different places of a loop body twice that size, and every warp
 * streams the body through the caches once per step.  Here the same sequence of evaluations runs as a small state machine:
 * a phase either measures ONE quarter-pel position (SAD or SATD: predictors, start point, half / quarter-pel refine) or
 * THREE full-pel positions around the current best (hexagon, square), and every phase goes through the same two pieces
 * of code.  The phase sequence is warp-uniform (phases nobody needs are skipped by vote); what a group does with the costs is
 * predicated, exactly as above.  Same arithmetic, same order of comparisons => same results (the parity suite runs both).
 * ------------------------------------------------------------------------------------------ */
enum
{
    PH_MVP0 = 0, PH_MVP1, PH_MVP2, PH_MVP3, PH_PRE, PH_SUB, PH_ZERO,           /* one quarter-pel position */
    PH_HEXA, PH_HEXB, PH_HEXI, PH_SQA, PH_SQB, PH_SQC,                        /* three full-pel positions */
    PH_HP1, PH_HP2, PH_HP3, PH_HP4, PH_RE, PH_QP1, PH_QP2, PH_QP3, PH_QP4,     /* one quarter-pel position */
    PH_DONE
};

template <typename Ctx>
__device__ __forceinline__ int searchBlockCompact(Ctx& m, int cand0, int cand1, int cand2, int cand3, bool valid0, bool valid1, bool valid2,
                                                  bool valid3, bool bidir, MV2 mvmin, MV2 mvmax, MV2& out, int& skipCostOut)
{
    const int merange = 16;
    const MV2 qmin = { mvmin.x << 2, mvmin.y << 2 }, qmax = { mvmax.x << 2, mvmax.y << 2 };
    /* identical predictors cost the same: each distinct one is measured once (slicetype.cpp:4158-4167 measures them all) */
    const int dup1 = (valid0 && cand0 == cand1) ? 0 : -1;
    const int dup2 = (valid0 && cand0 == cand2) ? 0 : (valid1 && cand1 == cand2) ? 1 : -1;
    const int dup3 = (valid0 && cand0 == cand3) ? 0 : (valid1 && cand1 == cand3) ? 1 : (valid2 && cand2 == cand3) ? 2 : -1;
    const bool need0 = valid0, need1 = valid1 && dup1 < 0, need2 = valid2 && dup2 < 0, need3 = valid3 && dup3 < 0;
    int pc0 = 0, pc1 = 0, pc2 = 0, pc3 = 0;       /* predictor costs */
    MV2 pmv = { 0, 0 }, bmv = { 0, 0 };
    int bprecost = 0, bcost = 0, dir = 0, iter = (merange >> 1) - 1, sqdir = 0, bdir = 0, zeroCost = 0;
    bool sub = false, nz = false, go = false, doSub = false;

    int phase = __any_sync(LA_FULL, need0) ? PH_MVP0 : __any_sync(LA_FULL, need1) ? PH_MVP1 : __any_sync(LA_FULL, need2) ? PH_MVP2 :
                __any_sync(LA_FULL, need3) ? PH_MVP3 : PH_PRE;
    bool first = true;      /* the predictor choice runs once, when the loop reaches PH_PRE */
#define LA_YOK(dy) ((bmv.y + (dy) >= mvmin.y) & (bmv.y + (dy) <= mvmax.y))
#define LA_INRANGE() (bmv.x >= mvmin.x && bmv.x <= mvmax.x && bmv.y >= mvmin.y && bmv.y <= mvmax.y)
#pragma unroll 1
    while (phase != PH_DONE)
    {
        if (phase == PH_PRE && first)
        {
            /* predictor choice, slicetype.cpp:4158-4168, in candidate order with the first minimum kept */
            first = false;
            if (dup1 == 0) pc1 = pc0;
            if (dup2 == 0) pc2 = pc0; else if (dup2 == 1) pc2 = pc1;
            if (dup3 == 0) pc3 = pc0; else if (dup3 == 1) pc3 = pc1; else if (dup3 == 2) pc3 = pc2;
            int mvpcost = LA_COST_MAX, skipCost = 0x7fffffff, mvpPacked = 0;
            if (valid0) { if (pc0 < mvpcost) { mvpcost = pc0; mvpPacked = cand0; } if (!mvpPacked && bidir) skipCost = pc0; }
            if (valid1) { if (pc1 < mvpcost) { mvpcost = pc1; mvpPacked = cand1; } if (!mvpPacked && bidir) skipCost = pc1; }
            if (valid2) { if (pc2 < mvpcost) { mvpcost = pc2; mvpPacked = cand2; } if (!mvpPacked && bidir) skipCost = pc2; }
            if (valid3) { if (pc3 < mvpcost) { mvpcost = pc3; mvpPacked = cand3; } if (!mvpPacked && bidir) skipCost = pc3; }
            skipCostOut = skipCost;
            const MV2 qmvp = unpackMv(mvpPacked);
            m.mvpx = qmvp.x; m.mvpy = qmvp.y;
            pmv.x = max(min(qmvp.x, qmax.x), qmin.x); pmv.y = max(min(qmvp.y, qmax.y), qmin.y);
            sub = ((pmv.x & 3) | (pmv.y & 3)) != 0;
            nz = (pmv.x | pmv.y) != 0;
        }
        if (phase >= PH_HEXA && phase <= PH_SQC)
        {
            /* ---- three full-pel SADs around bmv ---- */
            int pk;
            if (phase == PH_HEXA) pk = LA_PK3(-2, 0, -1, 2, 1, 2);
            else if (phase == PH_HEXB) pk = LA_PK3(2, 0, 1, -2, -1, -2);
            else if (phase == PH_HEXI)
                pk = LA_PK3(c_hex2[dir][0], c_hex2[dir][1], c_hex2[dir + 1][0], c_hex2[dir + 1][1], c_hex2[dir + 2][0], c_hex2[dir + 2][1]);
            else if (phase == PH_SQA) pk = LA_PK3(0, -1, 0, 1, -1, 0);
            else if (phase == PH_SQB) pk = LA_PK3(1, 0, -1, -1, -1, 1);
            else pk = LA_PK3(1, -1, 1, 1, 0, 0);        /* the third position of the last square triple is not used */
            const int x0 = (pk & 15) - 8, y0 = ((pk >> 4) & 15) - 8, x1 = ((pk >> 8) & 15) - 8, y1 = ((pk >> 12) & 15) - 8,
                      x2 = ((pk >> 16) & 15) - 8, y2 = ((pk >> 20) & 15) - 8;
            const int3 p = m.sad3Fpel(bmv.x, bmv.y, pk);
            const int c0 = p.x + m.mvc((bmv.x + x0) << 2, (bmv.y + y0) << 2);
            const int c1 = p.y + m.mvc((bmv.x + x1) << 2, (bmv.y + y1) << 2);
            const int c2 = p.z + m.mvc((bmv.x + x2) << 2, (bmv.y + y2) << 2);
            if (phase <= PH_HEXI)
            {
                /* hexagon, motion.cpp:892-946: bcost carries the winning direction in its low three bits */
                if (phase == PH_HEXA)
                {
                    if (LA_YOK(0)) bcost = min(bcost, (c0 << 3) + 2);
                    if (LA_YOK(2)) { bcost = min(bcost, (c1 << 3) + 3); bcost = min(bcost, (c2 << 3) + 4); }
                    phase = PH_HEXB;
                }
                else
                {
                    if (phase == PH_HEXB)
                    {
                        if (LA_YOK(0)) bcost = min(bcost, (c0 << 3) + 5);
                        if (LA_YOK(-2)) { bcost = min(bcost, (c1 << 3) + 6); bcost = min(bcost, (c2 << 3) + 7); }
                        go = (bcost & 7) != 0;
                        if (go)
                        {
                            dir = (bcost & 7) - 2;
                            go = LA_YOK(c_hex2[dir + 1][1]);
                            if (go) { bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1]; }
                        }
                        go = go && iter > 0 && LA_INRANGE();
                    }
                    else if (go)
                    {
                        /* half hexagon; groups that already stopped measured harmlessly and ignore the result */
                        bcost &= ~7;
                        if (LA_YOK(y0)) bcost = min(bcost, (c0 << 3) + 1);
                        if (LA_YOK(y1)) bcost = min(bcost, (c1 << 3) + 2);
                        if (LA_YOK(y2)) bcost = min(bcost, (c2 << 3) + 3);
                        if (!(bcost & 7))
                            go = false;
                        else
                        {
                            dir += (bcost & 7) - 2;
                            dir = c_mod6m1[dir + 1];
                            bmv.x += c_hex2[dir + 1][0]; bmv.y += c_hex2[dir + 1][1];
                            iter--;
                            go = iter > 0 && LA_INRANGE();
                        }
                    }
                    if (__any_sync(LA_FULL, go)) phase = PH_HEXI;
                    else { phase = PH_SQA; bcost >>= 3; sqdir = 0; }
                }
            }
            else if (phase == PH_SQA)
            {
                /* square refine, motion.cpp:950-967 */
                if (LA_YOK(-1)) { if (c0 < bcost) { bcost = c0; sqdir = 1; } }
                if (LA_YOK(1))  { if (c1 < bcost) { bcost = c1; sqdir = 2; } }
                if (c2 < bcost) { bcost = c2; sqdir = 3; }
                phase = PH_SQB;
            }
            else if (phase == PH_SQB)
            {
                if (c0 < bcost) { bcost = c0; sqdir = 4; }
                if (LA_YOK(-1)) { if (c1 < bcost) { bcost = c1; sqdir = 5; } }
                if (LA_YOK(1))  { if (c2 < bcost) { bcost = c2; sqdir = 6; } }
                phase = PH_SQC;
            }
            else
            {
                if (LA_YOK(-1)) { if (c0 < bcost) { bcost = c0; sqdir = 7; } }
                if (LA_YOK(1))  { if (c1 < bcost) { bcost = c1; sqdir = 8; } }
                bmv.x += c_square1[sqdir][0]; bmv.y += c_square1[sqdir][1];
                if (bprecost < bcost) { bmv = pmv; bcost = bprecost; }
                else { bmv.x <<= 2; bmv.y <<= 2; }
                /* zero residual at the start point: no subpel, cost = mvcost (motion.cpp:1490-1495) */
                doSub = bcost != 0;
                zeroCost = m.mvc(bmv.x, bmv.y);
                bdir = 0;
                phase = __any_sync(LA_FULL, doSub) ? PH_HP1 : PH_DONE;
            }
        }
        else
        {
            /* ---- one quarter-pel position: SAD or SATD of the motion-compensated block ---- */
            int qx, qy;
            if (phase <= PH_MVP3)
            {
                const int ci = phase == PH_MVP0 ? cand0 : phase == PH_MVP1 ? cand1 : phase == PH_MVP2 ? cand2 : cand3;
                const bool ni = phase == PH_MVP0 ? need0 : phase == PH_MVP1 ? need1 : phase == PH_MVP2 ? need2 : need3;
                const MV2 c = unpackMv(ni ? ci : 0);
                qx = c.x; qy = c.y;
            }
            else if (phase == PH_PRE) { qx = pmv.x; qy = pmv.y; }
            else if (phase == PH_SUB) { qx = ((pmv.x + 2) >> 2) << 2; qy = ((pmv.y + 2) >> 2) << 2; }
            else if (phase == PH_ZERO) { qx = 0; qy = 0; }
            else if (phase <= PH_HP4) { qx = bmv.x + c_square1[phase - PH_HP1 + 1][0] * 2; qy = bmv.y + c_square1[phase - PH_HP1 + 1][1] * 2; }
            else if (phase == PH_RE) { qx = bmv.x; qy = bmv.y; }
            else { qx = bmv.x + c_square1[phase - PH_QP1 + 1][0]; qy = bmv.y + c_square1[phase - PH_QP1 + 1][1]; }
            const bool satd = phase <= PH_MVP3 || phase >= PH_RE;      /* uniform */
            const int raw = m.qpelCost(qx, qy, satd);
            if (phase <= PH_MVP3)
            {
                if (phase == PH_MVP0) pc0 = raw; else if (phase == PH_MVP1) pc1 = raw; else if (phase == PH_MVP2) pc2 = raw; else pc3 = raw;
                phase = (phase < PH_MVP1 && __any_sync(LA_FULL, need1)) ? PH_MVP1 :
                        (phase < PH_MVP2 && __any_sync(LA_FULL, need2)) ? PH_MVP2 :
                        (phase < PH_MVP3 && __any_sync(LA_FULL, need3)) ? PH_MVP3 : PH_PRE;
            }
            else if (phase == PH_PRE)
            {
                /* motion.cpp:796-803: the start point is measured without its mv cost */
                bprecost = raw;
                bmv.x = (pmv.x + 2) >> 2; bmv.y = (pmv.y + 2) >> 2;
                bcost = bprecost;
                phase = __any_sync(LA_FULL, sub) ? PH_SUB : __any_sync(LA_FULL, nz) ? PH_ZERO : PH_HEXA;
                if (phase == PH_HEXA) bcost <<= 3;
            }
            else if (phase == PH_SUB)
            {
                if (sub) bcost = raw + m.mvc(qx, qy);
                phase = __any_sync(LA_FULL, nz) ? PH_ZERO : PH_HEXA;
                if (phase == PH_HEXA) bcost <<= 3;
            }
            else if (phase == PH_ZERO)
            {
                const int cost = raw + m.mvc(0, 0);
                if (nz && cost < bcost)
                {
                    bcost = cost;
                    bmv.x = 0;
                    bmv.y = max(min(0, mvmax.y), mvmin.y);
                }
                bcost <<= 3;
                phase = PH_HEXA;
            }
            else
            {
                /* lowres subpel, motion.cpp:1496-1528: 4 half-pel SADs, re-measure with SATD, 4 quarter-pel SATDs */
                const int cost = raw + m.mvc(qx, qy);
                if (phase == PH_RE)
                {
                    if (doSub) bcost = cost;
                    bdir = 0;
                }
                else
                {
                    const bool ok = doSub && !((qy < qmin.y) | (qy > qmax.y));
                    const int i = phase <= PH_HP4 ? phase - PH_HP1 + 1 : phase - PH_QP1 + 1;
                    if (ok && cost < bcost) { bcost = cost; bdir = i; }
                    if (phase == PH_HP4) { bmv.x += c_square1[bdir][0] * 2; bmv.y += c_square1[bdir][1] * 2; }
                    if (phase == PH_QP4) { bmv.x += c_square1[bdir][0]; bmv.y += c_square1[bdir][1]; }
                }
                phase++;        /* PH_QP4 + 1 == PH_DONE */
            }
        }
    }
#undef LA_YOK
#undef LA_INRANGE
    if (!doSub) bcost = zeroCost;
    out = bmv;
    return bcost;
}


template <typename P, int LPB>
__global__ void __launch_bounds__(32, LPB == 8 ? LA_SEARCH_MIN_CTAS : LA_SEARCH4_MIN_CTAS)
search_kernel(Geom g, const SearchJobDev<P>* __restrict__ jobs, int nstrips, int njobs, const unsigned short* __restrict__ mvcost,
              int* ticketCounter, int* progress, unsigned long long* executed, int oneShot)
{
    /* LPB lanes per block: 8 (one row per lane, 4 block rows per strip) or 4 (two rows per lane, 8 block rows per strip) */
    const int STRIP_ROWS = 32 / LPB;
    const int lane = threadIdx.x;
    const int grp = lane / LPB, r = lane % LPB;
    const int bw = g.bw, bh = g.bh;
    typename MeSel<P, LPB>::T m;
    m.mvcost = mvcost; m.init(lane);
    m.rb.planeSize = g.planeSize; m.rb.tpr = g.tpr;
    /* A warp is a worker: it takes strips from the ticket counter until none are left.  The grid may hold fewer
     * warps than strips (engine.cu sizes it); a warp that finishes a strip picks up the next one, whose
     * predecessor is by then well under way, instead of a fresh warp sitting resident while it waits its turn.
     * Strip-major tickets: every job's strip 0 first, then every strip 1, ...  A strip's predecessor always holds
     * a lower ticket, i.e. it is running or done, so polling on it cannot deadlock. */
    for (;;)
    {
    int ticket = 0;
    if (lane == 0) ticket = atomicAdd(ticketCounter, 1);
    ticket = __shfl_sync(LA_FULL, ticket, 0);
    if (ticket >= nstrips * njobs) return;
    const int strip = ticket / njobs, job = ticket % njobs;
    const SearchJobDev<P> J = jobs[job];
    if (J.cond && __ldcg(J.cond) == 0) continue;    /* the variant this job would compute is not needed */
    /* strip 0 owns the lowest ticket of its job and every other strip waits on its progress chain */
    if (strip == 0 && lane == 0) { atomicExch(J.flagOut, 0); atomicAdd(executed, 1ull); }
    const int rowsInStrip = min(STRIP_ROWS, bh - strip * STRIP_ROWS);
    const bool rowOk = grp < rowsInStrip;
    const int cuY = rowOk ? bh - 1 - strip * STRIP_ROWS - grp : 0;
    /* the first row a search visits takes no predictors from below: the frame's bottom row, or with cooperative
     * slices (slicetype.cpp:3957-3968) the bottom row of each slice, the last slice running to the frame's end */
    bool lastRow = cuY == bh - 1;
    if (J.sliced && g.rowsPerSlice > 0)
    {
        const int ns = bh / g.rowsPerSlice;
        const int si = min(cuY / g.rowsPerSlice, ns - 1);
        lastRow = cuY == (si == ns - 1 ? bh - 1 : (si + 1) * g.rowsPerSlice - 1);
    }
    const int steps = bw + 2 * (rowsInStrip - 1);
    int* myProgress = progress + job * nstrips + strip;
    const int* belowProgress = myProgress - 1;
    const int* belowRow = J.mvOut + min(cuY + 1, bh - 1) * bw;    /* group 0: last row of the strip below (global) */
    m.rb.plane0 = J.ref0;
    int h0 = 0, h1 = 0, h2 = 0;     /* this group's results of the last three steps (packed MVs) */
    int seen = 0;                   /* progress of the strip below as last observed (warp-uniform) */
    bool flagSet = false;

    for (int s = 0; s < steps; s++)
    {
        const int kRaw = s - 2 * grp;
        const bool act = rowOk && kRaw >= 0 && kRaw < bw;
        const int k = min(max(kRaw, 0), bw - 1);     /* idle groups shadow a valid block and write nothing */
        const int cuX = bw - 1 - k;
        const int cu = cuX + cuY * bw;
        m.rb.X0 = g.mx + 8 * cuX; m.rb.Y0 = g.my + 8 * cuY;
        m.loadFenc(J.fenc0);

        /* reverse-order MV predictors (slicetype.cpp:4131-4141): right, below, below-left, below-right */
        int cand[4]; bool valid[4];
        /* idle groups get no predictors: they must never follow an MV that is not there yet */
        valid[0] = act && cuX < bw - 1;
        valid[1] = act && !lastRow; valid[2] = act && !lastRow && cuX > 0; valid[3] = act && !lastRow && cuX < bw - 1;
        cand[0] = h0;
        /* the row below lives in the group below: it finished columns cuX-1, cuX, cuX+1 one, two and three steps ago */
        cand[2] = __shfl_up_sync(LA_FULL, h0, LPB);
        cand[1] = __shfl_up_sync(LA_FULL, h1, LPB);
        cand[3] = __shfl_up_sync(LA_FULL, h2, LPB);
        if (strip > 0 && s < bw)
        {
            /* row below group 0 belongs to the strip below: wait until it has finished column cuX-1 */
            const int need = min(bw, s + 2);
            if (seen < need)
            {
                if (lane == 0)
                    while ((seen = ldRelaxed(belowProgress)) < need) __nanosleep(400);
                seen = __shfl_sync(LA_FULL, seen, 0);
            }
            if (grp == 0)
            {
                cand[1] = __ldcg(belowRow + cuX);
                cand[2] = __ldcg(belowRow + max(cuX - 1, 0));
                cand[3] = __ldcg(belowRow + min(cuX + 1, bw - 1));
            }
        }
        const MV2 mvmin = { -cuX * 8 - 8, -cuY * 8 - 8 };
        const MV2 mvmax = { (bw - cuX - 1) * 8 + 8, (bh - cuY - 1) * 8 + 8 };
        MV2 best;
        int skipCost = 0x7fffffff;
#if LA_ME_COMPACT
        int fencCost = searchBlockCompact(m, cand[0], cand[1], cand[2], cand[3], valid[0], valid[1], valid[2], valid[3], J.bidir != 0,
                                          mvmin, mvmax, best, skipCost);
#else
        MV2 mvp = { 0, 0 };
#if LA_MVP_LOOP
        {
            /* the same, as a real loop: one copy of the quarter-pel SATD in the instruction stream instead of four */
            int mvpcost = LA_COST_MAX;
            int c0 = 0, c1 = 0, c2 = 0;
#pragma unroll 1
            for (int i = 0; i < 4; i++)
            {
                const int ci = i == 0 ? cand[0] : i == 1 ? cand[1] : i == 2 ? cand[2] : cand[3];
                const bool vi = i == 0 ? valid[0] : i == 1 ? valid[1] : i == 2 ? valid[2] : valid[3];
                int dup = -1;
                if (i > 0 && valid[0] && cand[0] == ci) dup = 0;
                else if (i > 1 && valid[1] && cand[1] == ci) dup = 1;
                else if (i > 2 && valid[2] && cand[2] == ci) dup = 2;
                const bool need = vi && dup < 0;
                int cost = 0;
                if (__any_sync(LA_FULL, need))
                {
                    const MV2 zero = { 0, 0 };
                    const MV2 c = need ? unpackMv(ci) : zero;
                    cost = m.qpelSatd(c.x, c.y);
                }
                if (dup == 0) cost = c0; else if (dup == 1) cost = c1; else if (dup == 2) cost = c2;
                if (i == 0) c0 = cost; else if (i == 1) c1 = cost; else if (i == 2) c2 = cost;
                if (vi)
                {
                    if (cost < mvpcost) { mvpcost = cost; mvp = unpackMv(ci); }
                    if (!(mvp.x | mvp.y) && J.bidir)
                        skipCost = cost;
                }
            }
        }
#else
        {
            int mvpcost = LA_COST_MAX;
            int costs[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
            {
                /* identical predictors cost the same: measure each distinct one once */
                int dup = -1;
#pragma unroll
                for (int q = 0; q < i; q++)
                    if (valid[q] && cand[q] == cand[i] && dup < 0) dup = q;
                const bool need = valid[i] && dup < 0;
                int cost = 0;
                if (__any_sync(LA_FULL, need))
                {
                    const MV2 zero = { 0, 0 };
                    const MV2 c = need ? unpackMv(cand[i]) : zero;
                    cost = m.qpelSatd(c.x, c.y);
                }
#pragma unroll
                for (int q = 0; q < i; q++)
                    if (dup == q) cost = costs[q];
                costs[i] = cost;
                if (valid[i])
                {
                    if (cost < mvpcost) { mvpcost = cost; mvp = unpackMv(cand[i]); }
                    if (!(mvp.x | mvp.y) && J.bidir)
                        skipCost = cost;
                }
            }
        }
#endif
        int fencCost = motionEstimate(m, mvmin, mvmax, mvp, best);
#endif
        bool skipped = false;
        if (skipCost < 64 && skipCost < fencCost && J.bidir)
        {
            fencCost = skipCost;
            best.x = best.y = 0;
            skipped = true;
        }
        const int packed = packMv(best);
        h2 = h1; h1 = h0; h0 = packed;
        if (r == 0 && act)
        {
            if (skipped && !flagSet) { atomicOr(J.flagOut, 1); flagSet = true; }    /* once per strip and row */
            __stcg(J.mvOut + cu, packed);
            J.costOut[cu] = fencCost;
            if (grp == rowsInStrip - 1)
                stRelease(myProgress, k + 1);      /* orders this thread's MV store before the counter */
        }
    }
    __syncwarp();
    /* one ticket per CTA (the grid then holds one CTA per ticket): a warp slot is given back after every strip, so pending
     * CTAs of higher-priority streams (pre-lookahead of the next frames, cuTree) get onto the SMs while a long search
     * launch is running instead of behind it */
    if (oneShot) return;
    }
}

/* ------------------------------------------------------------------------------------------
 * --hme (x265_param::bEnableHME): hierarchical motion estimation, the two levels the lookahead runs.
 *
 *   lowerres_kernel   Lowres::init's second downscale (lowres.cpp:378-388): lowresPlane[0] -> the four 1/16-resolution
 *                     planes, tiled like the lowres planes, in a buffer of their own geometry (Geom `g4`: w / 2 x h / 2
 *                     samples, half the margins); extend_border_kernel<P>(g4, ...) then fills their margins
 *   search_hme_kernel level 0 = the search of search_kernel on those planes over the m_4x4 block grid (slicetype.cpp:
 *                     4040-4048), level 1 = the lowres search with one more predictor, twice the level-0 vector of the
 *                     block's parent (:4142-4145).  Per level: search method dia / hex / umh and range (:4094-4096,
 *                     4170; motion.cpp:842) -- la_me_generic.cuh
 *
 * Same strips, tickets and progress chain as search_kernel.  Inside a wavefront step the four 8-lane groups of the warp
 * run their block's search INDEPENDENTLY (shuffles name the group's lanes only; idle groups skip the step): UMH is
 * data-dependent sequential control flow per block, not something four blocks can do in lockstep.  The warp reconverges
 * at the end of every step for the predictor hand-over.  This is a correctness-first path: the default configuration
 * (no --hme) never launches it.
 * ------------------------------------------------------------------------------------------ */
template <typename P>
__global__ void __launch_bounds__(256) lowerres_kernel(Geom g, Geom g4, const P* __restrict__ planes, P* __restrict__ planes4)
{
    /* one thread per sample of the tile-aligned interior: columns past w4 (w4 is a multiple of 4, tiles are 8 wide)
     * replicate the last sample like the margin the border kernel would have written there */
    const int wT = (g4.w + 7) & ~7;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)wT * g4.h) return;
    const int y = (int)(idx / wT), xo = (int)(idx % wT);
    const int x = min(xo, g4.w - 1);
    int a[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++)
            a[j][i] = planes[tileOff(g.mx + 2 * x + i, g.my + 2 * y + j, g.tpr)];
#define LA_F4(p, q, r, s) ((((p + q + 1) >> 1) + ((r + s + 1) >> 1) + 1) >> 1)       /* pixel.cpp:605-628 */
    const long long o = tileOff(g4.mx + xo, g4.my + y, g4.tpr);
    planes4[o]                     = (P)LA_F4(a[0][0], a[1][0], a[0][1], a[1][1]);
    planes4[o + g4.planeSize]      = (P)LA_F4(a[0][1], a[1][1], a[0][2], a[1][2]);
    planes4[o + 2 * g4.planeSize]  = (P)LA_F4(a[1][0], a[2][0], a[1][1], a[2][1]);
    planes4[o + 3 * g4.planeSize]  = (P)LA_F4(a[1][1], a[2][1], a[1][2], a[2][2]);
#undef LA_F4
}

template <typename P>
struct SearchJobHme
{
    const P* fenc0;         /* tiled plane-0 buffer of this level of the frame being searched */
    const P* ref0;          /* ... of the reference (level 1: possibly the weighted copy; level 0: never weighted, slicetype.cpp:4083) */
    int*     mvOut;         /* this level's packed MVs */
    int*     costOut;
    int*     flagOut;       /* the job's skip flag, shared by both levels: level 0 clears it, either level sets it */
    const int* cond;        /* as SearchJobDev::cond */
    const int* hmeMv;       /* level 1: the level-0 results of the same (frame, list, distance); level 0: NULL */
    const int* hmeCost;
    int      bidir;
    int      pad;
};

/* candidate evaluators of ONE 8-lane group (see la_me_generic.cuh for the interface).  Loads are clamped to the plane
 * buffer: the extra predictor of level 1 comes from another block's level-0 search and is followed unchecked by the
 * reference (it reads whatever lies outside its margins there); vectors inside the margins are not affected */
template <typename P>
struct MeCtxG
{
    Row<P> fenc;
    const P* plane0; long long planeSize; int tpr;
    int X0, Y0, maxX, maxY;
    const unsigned short* mvcost;   /* centre */
    int mvpx, mvpy;
    int r;
    unsigned mask;

    __device__ __forceinline__ Row<P> row(int pl, int X, int Y) const
    {
        return loadRowT(plane0 + pl * planeSize, tpr, min(max(X, 0), maxX), min(max(Y, 0), maxY));
    }
    __device__ __forceinline__ int mvc(int qx, int qy) const
    {
        return (int)(unsigned short)(__ldg(mvcost + (qx - mvpx)) + __ldg(mvcost + (qy - mvpy)));
    }
    __device__ __forceinline__ int sadFpel(int x, int y) const { return groupSumM(sadRow(fenc, row(0, X0 + x, Y0 + y + r)), mask); }
    __device__ __forceinline__ Row<P> mc(int qx, int qy) const      /* lowres.h:71-96; (qx, qy) is uniform inside the group */
    {
        const int hA = (qy & 2) | ((qx & 2) >> 1);
        const Row<P> A = row(hA, X0 + (qx >> 2), Y0 + (qy >> 2) + r);
        if (!((qx | qy) & 1)) return A;
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hB = (qy2 & 2) | ((qx2 & 2) >> 1);
        return avgRow(A, row(hB, X0 + (qx2 >> 2), Y0 + (qy2 >> 2) + r));
    }
    __device__ __forceinline__ int qpelSad(int qx, int qy) const { return groupSumM(sadRow(fenc, mc(qx, qy)), mask); }
    __device__ __forceinline__ int qpelSatd(int qx, int qy) const { return groupSatdRowsM<P>(fenc, mc(qx, qy), mask); }
};

template <typename P>
__global__ void __launch_bounds__(32, 16)
search_hme_kernel(Geom g /* geometry of this level */, const SearchJobHme<P>* __restrict__ jobs, int nstrips, int njobs,
                  const unsigned short* __restrict__ mvcost, int* ticketCounter, int* progress, unsigned long long* executed,
                  int level, int method, int merange)
{
    const int STRIP_ROWS = 4;
    const int lane = threadIdx.x;
    const int grp = lane >> 3, r = lane & 7;
    const int bw = g.bw, bh = g.bh;
    MeCtxG<P> m;
    m.mvcost = mvcost; m.r = r; m.mask = 0xffu << (grp * 8);
    m.planeSize = g.planeSize; m.tpr = g.tpr;
    m.maxX = g.tpr * 8 - 16; m.maxY = g.planeLines - 1;
    for (;;)
    {
        int ticket = 0;
        if (lane == 0) ticket = atomicAdd(ticketCounter, 1);
        ticket = __shfl_sync(LA_FULL, ticket, 0);
        if (ticket >= nstrips * njobs) return;
        const int strip = ticket / njobs, job = ticket % njobs;
        const SearchJobHme<P> J = jobs[job];
        if (J.cond && __ldcg(J.cond) == 0) continue;
        if (strip == 0 && lane == 0)
        {
            if (level == 0) atomicExch(J.flagOut, 0);
            else atomicAdd(executed, 1ull);
        }
        const int rowsInStrip = min(STRIP_ROWS, bh - strip * STRIP_ROWS);
        const bool rowOk = grp < rowsInStrip;
        const int cuY = rowOk ? bh - 1 - strip * STRIP_ROWS - grp : 0;
        const bool lastRow = cuY == bh - 1;         /* no cooperative slices with --hme (refused at create) */
        const int steps = bw + 2 * (rowsInStrip - 1);
        int* myProgress = progress + job * nstrips + strip;
        const int* belowProgress = myProgress - 1;
        const int* belowRow = J.mvOut + min(cuY + 1, bh - 1) * bw;
        m.plane0 = J.ref0;
        int h0 = 0, h1 = 0, h2 = 0;
        int seen = 0;
        bool flagSet = false;

        for (int s = 0; s < steps; s++)
        {
            const int kRaw = s - 2 * grp;
            const bool act = rowOk && kRaw >= 0 && kRaw < bw;
            const int k = min(max(kRaw, 0), bw - 1);
            const int cuX = bw - 1 - k;
            const int cu = cuX + cuY * bw;
            /* reverse-order MV predictors (slicetype.cpp:4131-4141): right, below, below-left, below-right */
            int cand[5]; bool valid[5];
            valid[0] = act && cuX < bw - 1;
            valid[1] = act && !lastRow; valid[2] = act && !lastRow && cuX > 0; valid[3] = act && !lastRow && cuX < bw - 1;
            cand[0] = h0;
            cand[2] = __shfl_up_sync(LA_FULL, h0, 8);
            cand[1] = __shfl_up_sync(LA_FULL, h1, 8);
            cand[3] = __shfl_up_sync(LA_FULL, h2, 8);
            if (strip > 0 && s < bw)
            {
                const int need = min(bw, s + 2);
                if (seen < need)
                {
                    if (lane == 0)
                        while ((seen = ldRelaxed(belowProgress)) < need) __nanosleep(400);
                    seen = __shfl_sync(LA_FULL, seen, 0);
                }
                if (grp == 0)
                {
                    cand[1] = __ldcg(belowRow + cuX);
                    cand[2] = __ldcg(belowRow + max(cuX - 1, 0));
                    cand[3] = __ldcg(belowRow + min(cuX + 1, bw - 1));
                }
            }
            valid[4] = false; cand[4] = 0;
            if (act && J.hmeMv)
            {
                /* the level-0 block this block is read from, exactly as indexed at slicetype.cpp:4088 */
                const int cu4 = (cuX / 2) + (cuY / 2) * bw / 2;
                if (__ldg(J.hmeCost + cu4) > 0)
                {
                    const MV2 v = unpackMv(__ldg(J.hmeMv + cu4));
                    const MV2 v2 = { v.x * 2, v.y * 2 };
                    cand[4] = packMv(v2); valid[4] = true;
                }
            }
            int packed = 0;
            if (act)
            {   /* from here to the end of the branch only this group's lanes take part */
                m.X0 = g.mx + 8 * cuX; m.Y0 = g.my + 8 * cuY;
                m.fenc = loadRowAligned(J.fenc0, g.tpr, m.X0, m.Y0 + r);
                const MV2 mvmin = { -cuX * 8 - 8, -cuY * 8 - 8 };
                const MV2 mvmax = { (bw - cuX - 1) * 8 + 8, (bh - cuY - 1) * 8 + 8 };
                MV2 mvp = { 0, 0 };
                int skipCost = 0x7fffffff, mvpcost = LA_COST_MAX;
#pragma unroll 1
                for (int i = 0; i < 5; i++)
                {
                    const int ci = i == 0 ? cand[0] : i == 1 ? cand[1] : i == 2 ? cand[2] : i == 3 ? cand[3] : cand[4];
                    const bool vi = i == 0 ? valid[0] : i == 1 ? valid[1] : i == 2 ? valid[2] : i == 3 ? valid[3] : valid[4];
                    if (!vi) continue;
                    const MV2 c = unpackMv(ci);
                    const int cost = m.qpelSatd(c.x, c.y);
                    if (cost < mvpcost) { mvpcost = cost; mvp = c; }
                    if (!(mvp.x | mvp.y) && J.bidir)
                        skipCost = cost;
                }
                MV2 best;
                int fencCost = motionEstimateG(m, mvmin, mvmax, mvp, merange, method, best);
                bool skipped = false;
                if (skipCost < 64 && skipCost < fencCost && J.bidir)
                {
                    fencCost = skipCost;
                    best.x = best.y = 0;
                    skipped = true;
                }
                packed = packMv(best);
                if (r == 0)
                {
                    if (skipped && !flagSet) { atomicOr(J.flagOut, 1); flagSet = true; }
                    __stcg(J.mvOut + cu, packed);
                    J.costOut[cu] = fencCost;
                    if (grp == rowsInStrip - 1)
                        stRelease(myProgress, k + 1);
                }
            }
            __syncwarp();
            /* every group shifts at every step, idle or not: the row above reads (h0, h1, h2) as "columns cuX-1, cuX, cuX+1 of
             * the row below", which keeps sliding for two steps after that row has finished */
            h2 = h1; h1 = h0; h0 = packed;
        }
        __syncwarp();
    }
}

/* ------------------------------------------------------------------------------------------
 * K5: frame cost.  Cost half of estimateCUCost (slicetype.cpp:4187-4248) + the frame sums of
 * estimateFrameCost (:4050-4062).  P estimates need no pixels (cost_p_kernel, one thread per block); B estimates
 * evaluate the two bidir candidates with 8 lanes per block (cost_group_kernel).  Sums are integer atomics, so
 * the result does not depend on the order of accumulation.
 * ------------------------------------------------------------------------------------------ */
template <typename P>
struct CostJobDev
{
    const P* fenc0; const P* ref0; const P* ref1;   /* tiled plane-0 buffers of b, p0, p1 (ref1 = NULL: P estimate) */
    const int* mv0; const int* cost0; const int* mv1; const int* cost1;
    const int* intraCost; const int* invQ;
    unsigned short* lowresCosts; int* rowSatds; CostResultDev* result;
    const int* cond;        /* NULL, or a search's skip flag: run only if it is set (see SearchJobDev::cond) */
};

template <typename P>
__global__ void __launch_bounds__(256) cost_clear_kernel(Geom g, const CostJobDev<P>* __restrict__ jobs, unsigned long long* executed)
{
    const CostJobDev<P> J = jobs[blockIdx.x];
    if (J.cond && __ldcg(J.cond) == 0) return;
    if (threadIdx.x == 0) atomicAdd(executed, 1ull);
    for (int i = threadIdx.x; i < g.bh; i += blockDim.x) J.rowSatds[i] = 0;
    if (threadIdx.x == 0) { J.result->costEst = 0; J.result->costEstAq = 0; J.result->intraMbs = 0; J.result->reserved = 0; }
}

/* P estimate: no pixels, one thread per block (slicetype.cpp:4209-4218) */
template <typename P>
__global__ void __launch_bounds__(128) cost_p_kernel(Geom g, const CostJobDev<P>* __restrict__ jobs)
{
    __shared__ unsigned long long s_acc[2];
    __shared__ int s_intra;
    const CostJobDev<P> J = jobs[blockIdx.y];
    if (J.cond && __ldcg(J.cond) == 0) return;      /* uniform for the whole CTA */
    if (threadIdx.x < 2) s_acc[threadIdx.x] = 0;
    if (threadIdx.x == 2) s_intra = 0;
    __syncthreads();
    const int cu = blockIdx.x * 128 + threadIdx.x;
    if (cu < g.ncu)
    {
        int bcost = J.cost0[cu] + 4, listused = 1;      /* COST_MAX > any search cost */
        const int ic = J.intraCost[cu];
        if (ic < bcost) { bcost = ic; listused = 0; }
        const int cuX = cu % g.bw, cuY = cu / g.bw;
        const bool scored = (cuX > 0 && cuX < g.bw - 1 && cuY > 0 && cuY < g.bh - 1) || g.bw <= 2 || g.bh <= 2;
        const int bcostAq = (scored && J.invQ) ? ((bcost * J.invQ[cu] + 128) >> 8) : bcost;
        if (scored)
        {
            atomicAdd(&s_acc[0], (unsigned long long)bcost);
            atomicAdd(&s_acc[1], (unsigned long long)bcostAq);
            if (!listused) atomicAdd(&s_intra, 1);
        }
        atomicAdd(&J.rowSatds[cuY], bcostAq);
        J.lowresCosts[cu] = (unsigned short)(min(bcost, LA_LOWRES_COST_MASK) | (listused << LA_LOWRES_COST_SHIFT));
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (s_acc[0]) atomicAdd((unsigned long long*)&J.result->costEst, s_acc[0]);
        if (s_acc[1]) atomicAdd((unsigned long long*)&J.result->costEstAq, s_acc[1]);
        if (s_intra) atomicAdd(&J.result->intraMbs, s_intra);
    }
}

/* B estimates, grouped: every (p0, p1, b) of a batch that shares b and p1 -- i.e. the same source block, the same
 * list-1 motion compensation and the same co-located list-1 block -- is one group; a CTA loads those once and walks
 * the group's list-0 references.  (One launch per estimate re-fetched them up to bframes times: ncu showed the
 * load pipe and L1 as that kernel's limit.)  8 lanes per block, 16 blocks per CTA, warp-uniform. */
#define LA_COST_GROUP_MAX 16

template <typename P>
struct CostGroupDev
{
    const P* fenc0; const P* ref1;          /* tiled plane-0 buffers of b and p1 */
    const int* mv1; const int* cost1;       /* list-1 search of b towards p1 */
    const int* intraCost; const int* invQ;
    int n, pad;
    struct Member
    {
        const P* ref0; const int* mv0; const int* cost0;        /* p0 and the list-0 search of b towards it */
        unsigned short* lowresCosts; int* rowSatds; CostResultDev* result;
        const int* cond;
    } m[LA_COST_GROUP_MAX];
};

template <typename P>
__global__ void __launch_bounds__(128) cost_group_kernel(Geom g, const CostGroupDev<P>* __restrict__ groups)
{
    __shared__ unsigned long long s_acc[LA_COST_GROUP_MAX][2];
    const CostGroupDev<P>& G = groups[blockIdx.y];
    const int n = G.n;
    if (threadIdx.x < 2 * LA_COST_GROUP_MAX) s_acc[threadIdx.x >> 1][threadIdx.x & 1] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int cuRaw = blockIdx.x * 16 + grp;
    const bool act = cuRaw < g.ncu;
    const int cu = act ? cuRaw : g.ncu - 1;
    const int cuX = cu % g.bw, cuY = cu / g.bw;
    const int X0 = g.mx + 8 * cuX, Y0 = g.my + 8 * cuY;
    const bool scored = (cuX > 0 && cuX < g.bw - 1 && cuY > 0 && cuY < g.bh - 1) || g.bw <= 2 || g.bh <= 2;
    const int invQ = (scored && G.invQ) ? G.invQ[cu] : 256;
    const Row<P> fenc = loadRowAligned(G.fenc0, g.tpr, X0, Y0 + r);
    const int c1 = G.cost1[cu];
    const MV2 m1 = unpackMv(G.mv1[cu]);
    RefBlock<P> rb1 = { G.ref1, g.planeSize, g.tpr, X0, Y0 };
    const Row<P> mc1 = mcRow(rb1, m1.x, m1.y, r);
    const Row<P> co1 = loadRowAligned(G.ref1, g.tpr, X0, Y0 + r);
    for (int i = 0; i < n; i++)
    {
        const typename CostGroupDev<P>::Member& M = G.m[i];
        if (M.cond && __ldcg(M.cond) == 0) continue;    /* uniform for the whole CTA */
        int bcost = LA_COST_MAX, listused = 0;
        const int c0 = M.cost0[cu];
        if (c0 < bcost) { bcost = c0; listused = 1; }
        if (c1 < bcost) { bcost = c1; listused = 2; }
        const MV2 m0 = unpackMv(M.mv0[cu]);
        RefBlock<P> rb0 = { M.ref0, g.planeSize, g.tpr, X0, Y0 };
        /* avg(L0 MC, L1 MC), unweighted references (slicetype.cpp:4189-4200) */
        const Row<P> a = avgRow(mcRow(rb0, m0.x, m0.y, r), mc1);
        /* co-located average (:4201-4206) */
        const Row<P> b = avgRow(loadRowAligned(M.ref0, g.tpr, X0, Y0 + r), co1);
        int bicost = groupSatdRows(fenc, a);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        bicost = groupSatdRows(fenc, b);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        bcost += 4;
        const int bcostAq = scored ? ((bcost * invQ + 128) >> 8) : bcost;
        const bool lead = r == 0 && act;
        if (lead)
        {
            atomicAdd(&M.rowSatds[cuY], bcostAq);
            M.lowresCosts[cu] = (unsigned short)(min(bcost, LA_LOWRES_COST_MASK) | (listused << LA_LOWRES_COST_SHIFT));
        }
        /* frame sums over the interior blocks: the warp's four blocks first, then one shared atomic per warp */
        int v0 = (lead && scored) ? bcost : 0, v1 = (lead && scored) ? bcostAq : 0;
        v0 += __shfl_xor_sync(LA_FULL, v0, 8); v1 += __shfl_xor_sync(LA_FULL, v1, 8);
        v0 += __shfl_xor_sync(LA_FULL, v0, 16); v1 += __shfl_xor_sync(LA_FULL, v1, 16);
        if (lane == 0 && (v0 | v1))
        {
            atomicAdd(&s_acc[i][0], (unsigned long long)v0);
            atomicAdd(&s_acc[i][1], (unsigned long long)v1);
        }
    }
    __syncthreads();
    if (threadIdx.x < n)
    {
        const typename CostGroupDev<P>::Member& M = G.m[threadIdx.x];
        if (s_acc[threadIdx.x][0]) atomicAdd((unsigned long long*)&M.result->costEst, s_acc[threadIdx.x][0]);
        if (s_acc[threadIdx.x][1]) atomicAdd((unsigned long long*)&M.result->costEstAq, s_acc[threadIdx.x][1]);
    }
}


/* the same with FOUR lanes per block (two rows per lane, la_device.cuh): 32 blocks per CTA, half the group-uniform work
 * and half the shuffles per block -- the kernel is two SATDs and four row fetches per estimate, nothing else */
template <typename P>
__device__ __forceinline__ void mcRows2(const RefBlock<P>& rb, int qx, int qy, int r2, Row<P>& o0, Row<P>& o1)
{
    const int hA = (qy & 2) | ((qx & 2) >> 1);
    const P* pA = rb.plane0 + hA * rb.planeSize;
    const int xa = rb.X0 + (qx >> 2), ya = rb.Y0 + (qy >> 2) + r2;
    o0 = loadRowT(pA, rb.tpr, xa, ya); o1 = loadRowT(pA, rb.tpr, xa, ya + 1);
    if (__any_sync(LA_FULL, (qx | qy) & 1))
    {
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const P* pB = rb.plane0 + hB * rb.planeSize;
        const int xb = rb.X0 + (qx2 >> 2), yb = rb.Y0 + (qy2 >> 2) + r2;
        o0 = avgRow(o0, loadRowT(pB, rb.tpr, xb, yb));
        o1 = avgRow(o1, loadRowT(pB, rb.tpr, xb, yb + 1));
    }
}

template <typename P>
__global__ void __launch_bounds__(128) cost_group_kernel4(Geom g, const CostGroupDev<P>* __restrict__ groups)
{
    __shared__ unsigned long long s_acc[LA_COST_GROUP_MAX][2];
    const CostGroupDev<P>& G = groups[blockIdx.y];
    const int n = G.n;
    if (threadIdx.x < 2 * LA_COST_GROUP_MAX) s_acc[threadIdx.x >> 1][threadIdx.x & 1] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int grp = threadIdx.x >> 2, r2 = (threadIdx.x & 3) * 2;
    const int cuRaw = blockIdx.x * 32 + grp;
    const bool act = cuRaw < g.ncu;
    const int cu = act ? cuRaw : g.ncu - 1;
    const int cuX = cu % g.bw, cuY = cu / g.bw;
    const int X0 = g.mx + 8 * cuX, Y0 = g.my + 8 * cuY;
    const bool scored = (cuX > 0 && cuX < g.bw - 1 && cuY > 0 && cuY < g.bh - 1) || g.bw <= 2 || g.bh <= 2;
    const int invQ = (scored && G.invQ) ? G.invQ[cu] : 256;
    const Row<P> fa = loadRowAligned(G.fenc0, g.tpr, X0, Y0 + r2), fb = loadRowAligned(G.fenc0, g.tpr, X0, Y0 + r2 + 1);
    const int c1 = G.cost1[cu];
    const MV2 m1 = unpackMv(G.mv1[cu]);
    RefBlock<P> rb1 = { G.ref1, g.planeSize, g.tpr, X0, Y0 };
    Row<P> mc1a, mc1b;
    mcRows2(rb1, m1.x, m1.y, r2, mc1a, mc1b);
    const Row<P> co1a = loadRowAligned(G.ref1, g.tpr, X0, Y0 + r2), co1b = loadRowAligned(G.ref1, g.tpr, X0, Y0 + r2 + 1);
    for (int i = 0; i < n; i++)
    {
        const typename CostGroupDev<P>::Member& M = G.m[i];
        if (M.cond && __ldcg(M.cond) == 0) continue;    /* uniform for the whole CTA */
        int bcost = LA_COST_MAX, listused = 0;
        const int c0 = M.cost0[cu];
        if (c0 < bcost) { bcost = c0; listused = 1; }
        if (c1 < bcost) { bcost = c1; listused = 2; }
        const MV2 m0 = unpackMv(M.mv0[cu]);
        RefBlock<P> rb0 = { M.ref0, g.planeSize, g.tpr, X0, Y0 };
        /* avg(L0 MC, L1 MC), unweighted references (slicetype.cpp:4189-4200) */
        Row<P> a0, a1;
        mcRows2(rb0, m0.x, m0.y, r2, a0, a1);
        a0 = avgRow(a0, mc1a); a1 = avgRow(a1, mc1b);
        int bicost = group4SatdRows(fa, fb, a0, a1);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        /* co-located average (:4201-4206) */
        const Row<P> b0 = avgRow(loadRowAligned(M.ref0, g.tpr, X0, Y0 + r2), co1a);
        const Row<P> b1 = avgRow(loadRowAligned(M.ref0, g.tpr, X0, Y0 + r2 + 1), co1b);
        bicost = group4SatdRows(fa, fb, b0, b1);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        bcost += 4;
        const int bcostAq = scored ? ((bcost * invQ + 128) >> 8) : bcost;
        const bool lead = (threadIdx.x & 3) == 0 && act;
        if (lead)
        {
            atomicAdd(&M.rowSatds[cuY], bcostAq);
            M.lowresCosts[cu] = (unsigned short)(min(bcost, LA_LOWRES_COST_MASK) | (listused << LA_LOWRES_COST_SHIFT));
        }
        /* frame sums over the interior blocks: the warp's eight blocks first, then one shared atomic per warp */
        int v0 = (lead && scored) ? bcost : 0, v1 = (lead && scored) ? bcostAq : 0;
        v0 += __shfl_xor_sync(LA_FULL, v0, 4); v1 += __shfl_xor_sync(LA_FULL, v1, 4);
        v0 += __shfl_xor_sync(LA_FULL, v0, 8); v1 += __shfl_xor_sync(LA_FULL, v1, 8);
        v0 += __shfl_xor_sync(LA_FULL, v0, 16); v1 += __shfl_xor_sync(LA_FULL, v1, 16);
        if (lane == 0 && (v0 | v1))
        {
            atomicAdd(&s_acc[i][0], (unsigned long long)v0);
            atomicAdd(&s_acc[i][1], (unsigned long long)v1);
        }
    }
    __syncthreads();
    if (threadIdx.x < n)
    {
        const typename CostGroupDev<P>::Member& M = G.m[threadIdx.x];
        if (s_acc[threadIdx.x][0]) atomicAdd((unsigned long long*)&M.result->costEst, s_acc[threadIdx.x][0]);
        if (s_acc[threadIdx.x][1]) atomicAdd((unsigned long long*)&M.result->costEstAq, s_acc[threadIdx.x][1]);
    }
}


/* ------------------------------------------------------------------------------------------
 * --hist-scenecut: LookaheadTLD::collectPictureStatistics (slicetype.cpp:1441-1724) per frame.  8-bit only (the reference
 * indexes its 256 bins with the sample value).  Accumulators per slot: counts[16][3][256] (u32), sums[16][3] (u64),
 * varTot[3] (u64), zeroed before the first kernel; hist_finish_kernel turns them into x265cu_hist_stats in mapped host memory.
 * Segment index = wi * 4 + hi (picHistogram[wi][hi]).
 * ------------------------------------------------------------------------------------------ */
struct HistAcc
{
    unsigned counts[16][3][256];
    unsigned long long sums[16][3];
    unsigned long long varTot[3];
};

struct HistStatsDev     /* mirrors x265cu_hist_stats */
{
    unsigned histogram[4][4][3][256];
    unsigned char avgIntensitySeg[4][4][3];
    unsigned char avgIntensity[3];
    unsigned char pad0;
    unsigned short picAvgVariance[3];
    unsigned short pad1;
};

/* luma: histogram of the quarter-sampled picture, frame_lowres_core applied to lowresPlane[0] (lowres.cpp:35-51, 392-402),
 * computed on the fly from the tiled plane.  grid (16 segments, LA_HIST_CHUNKS row chunks) */
#define LA_HIST_CHUNKS 8
__global__ void __launch_bounds__(256) hist_luma_kernel(Geom g, const uint8_t* __restrict__ plane0, HistAcc* acc)
{
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_sum;
    const int seg = blockIdx.x, wi = seg >> 2, hi = seg & 3;
    const int qW = g.picW / 4, qH = g.picH / 4, segW = qW / 4, segH = qH / 4;
    const int w = segW + (wi == 3 ? qW - 4 * segW : 0), h = segH + (hi == 3 ? qH - 4 * segH : 0);
    const int x0 = wi * segW, y0 = hi * segH;
    s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int rows = (h + LA_HIST_CHUNKS - 1) / LA_HIST_CHUNKS;
    const int r0 = blockIdx.y * rows, r1 = min(h, r0 + rows);
    unsigned long long sum = 0;
    for (int r = r0; r < r1; r++)
        for (int c = threadIdx.x; c < w; c += 256)
        {
            const int X = g.mx + 2 * (x0 + c), Y = g.my + 2 * (y0 + r);
            const int a = plane0[tileOff(X, Y, g.tpr)], b = plane0[tileOff(X, Y + 1, g.tpr)];
            const int cc = plane0[tileOff(X + 1, Y, g.tpr)], d = plane0[tileOff(X + 1, Y + 1, g.tpr)];
            const int v = (((a + b + 1) >> 1) + ((cc + d + 1) >> 1) + 1) >> 1;
            atomicAdd(&s_hist[v], 1u);
            sum += (unsigned)v;
        }
    atomicAdd(&s_sum, sum);
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&acc->counts[seg][0][threadIdx.x], s_hist[threadIdx.x]);
    if (threadIdx.x == 0 && s_sum) atomicAdd(&acc->sums[seg][0], s_sum);
}

/* chroma: every 4th sample of every 4th row of the segment (calculateHistogram with dsFactor 4).  grid (16 segments, 2 planes) */
__global__ void __launch_bounds__(256) hist_chroma_kernel(Geom g, const uint8_t* __restrict__ u, const uint8_t* __restrict__ v, HistAcc* acc)
{
    __shared__ unsigned s_hist[256];
    __shared__ unsigned long long s_sum;
    const int seg = blockIdx.x, wi = seg >> 2, hi = seg & 3, pl = 1 + blockIdx.y;
    const uint8_t* src = pl == 1 ? u : v;
    const int segW = g.picW / 4, segH = g.picH / 4;
    const int w = (segW + (wi == 3 ? g.picW - 4 * segW : 0)) >> 1, h = (segH + (hi == 3 ? g.picH - 4 * segH : 0)) >> 1;
    const int x0 = (wi * segW) >> 1, y0 = (hi * segH) >> 1;
    s_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int nx = (w + 3) / 4, ny = (h + 3) / 4;
    unsigned long long sum = 0;
    for (int i = threadIdx.x; i < nx * ny; i += 256)
    {
        const int c = (i % nx) * 4, r = (i / nx) * 4;
        const int vv = src[(long long)(y0 + r) * g.srcPitchC + x0 + c];
        atomicAdd(&s_hist[vv], 1u);
        sum += (unsigned)vv;
    }
    atomicAdd(&s_sum, sum);
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&acc->counts[seg][pl][threadIdx.x], s_hist[threadIdx.x]);
    if (threadIdx.x == 0) atomicAdd(&acc->sums[seg][pl], s_sum);
}

/* computePictureStatistics (:1457-1544): one CTA per block row of a plane; the row's block variances are summed, divided by the
 * plane width and CUT to 16 bits before they are added up.  Blocks read the picture with replicate clamping (PicYuv's padding).
 * grid = luma rows of 8 + 2 x chroma rows of 4 */
__global__ void __launch_bounds__(256) hist_var_kernel(Geom g, const uint8_t* __restrict__ y, const uint8_t* __restrict__ u,
                                                        const uint8_t* __restrict__ v, HistAcc* acc)
{
    __shared__ unsigned long long s_row;
    const int lumaRows = (g.picH + 7) / 8, cH = g.picH >> 1, cW = g.picW >> 1, chromaRows = (cH + 3) / 4;
    int pl, row;
    if ((int)blockIdx.x < lumaRows) { pl = 0; row = blockIdx.x; }
    else { pl = 1 + ((int)blockIdx.x - lumaRows) / chromaRows; row = ((int)blockIdx.x - lumaRows) % chromaRows; }
    const uint8_t* src = pl == 0 ? y : pl == 1 ? u : v;
    const int pitch = pl == 0 ? g.srcPitch : g.srcPitchC;
    const int size = pl == 0 ? 8 : 4, shift = pl == 0 ? 6 : 4;
    const int planeW = pl == 0 ? g.picW : g.cW, planeH = pl == 0 ? g.picH : g.cH;       /* clamp limits */
    const int loopW = pl == 0 ? g.picW : cW;                                            /* maxCol / maxColChroma */
    if (threadIdx.x == 0) s_row = 0;
    __syncthreads();
    unsigned long long mine = 0;
    const int nblk = (loopW + size - 1) / size;
    for (int b = threadIdx.x; b < nblk; b += 256)
    {
        unsigned sum = 0, sqr = 0;
        for (int yy = 0; yy < size; yy++)
        {
            const int Y = min(row * size + yy, planeH - 1);
            for (int xx = 0; xx < size; xx++)
            {
                const unsigned px = src[(long long)Y * pitch + min(b * size + xx, planeW - 1)];
                sum += px; sqr += px * px;
            }
        }
        mine += sqr - (unsigned)(((unsigned long long)sum * sum) >> shift);
    }
    atomicAdd(&s_row, mine);
    __syncthreads();
    if (threadIdx.x == 0)
        atomicAdd(&acc->varTot[pl], (unsigned long long)(unsigned short)(s_row / (unsigned long long)loopW));
}

/* the reference's final arithmetic, quirks included (slicetype.cpp:1604-1607, 1630-1633, 1683, 1715-1719) */
__global__ void __launch_bounds__(256) hist_finish_kernel(Geom g, const HistAcc* __restrict__ acc, HistStatsDev* out)
{
    const unsigned W = g.picW, H = g.picH;
    for (int i = threadIdx.x; i < 16 * 3 * 256; i += 256)
    {
        const int seg = i / 768, pl = (i / 256) % 3, bin = i & 255;
        out->histogram[seg >> 2][seg & 3][pl][bin] = (1u + acc->counts[seg][pl][bin]) << 4;
    }
    if (threadIdx.x < 16)
    {
        const int seg = threadIdx.x, wi = seg >> 2, hi = seg & 3;
        {
            const unsigned qW = W / 4, qH = H / 4, segW = qW / 4, segH = qH / 4;
            const unsigned offW = wi == 3 ? qW - 4 * segW : 0, offH = hi == 3 ? qH - 4 * segH : 0;
            const unsigned long long sum = acc->sums[seg][0];
            out->avgIntensitySeg[wi][hi][0] = (unsigned char)((sum + (((segW + offW) * (segW + offH)) >> 1)) / ((segW + offW) * (segH + offH)));
        }
        const unsigned segW = W / 4, segH = H / 4;
        const unsigned offW = wi == 3 ? W - 4 * segW : 0, offH = hi == 3 ? H - 4 * segH : 0;
        for (int pl = 1; pl <= 2; pl++)
        {
            const unsigned long long sum = acc->sums[seg][pl] << 4;
            const unsigned den = pl == 1 ? ((segW + offW) * (segH + offH)) >> 2 : ((segW + offH) * (segH + offH)) >> 2;
            out->avgIntensitySeg[wi][hi][pl] = (unsigned char)((sum + (((segW + offW) * (segH + offH)) >> 3)) / den);
        }
    }
    if (threadIdx.x == 0)
    {
        unsigned long long s0 = 0, s1 = 0, s2 = 0;
        for (int seg = 0; seg < 16; seg++) { s0 += acc->sums[seg][0] << 4; s1 += acc->sums[seg][1] << 4; s2 += acc->sums[seg][2] << 4; }
        const unsigned long long area = (unsigned long long)W * H;
        out->avgIntensity[0] = (unsigned char)((s0 + (area >> 1)) / area);
        out->avgIntensity[1] = (unsigned char)((s1 + (area >> 3)) / (area >> 2));
        out->avgIntensity[2] = (unsigned char)((s2 + (area >> 3)) / (area >> 2));
        out->picAvgVariance[0] = (unsigned short)(acc->varTot[0] / (unsigned long long)H);
        out->picAvgVariance[1] = (unsigned short)(acc->varTot[1] / (unsigned long long)(H >> 1));
        out->picAvgVariance[2] = (unsigned short)(acc->varTot[2] / (unsigned long long)(H >> 1));
        out->pad0 = 0; out->pad1 = 0;
    }
}

/* ------------------------------------------------------------------------------------------
 * K6: weighted prediction.  weight_pp_c over whole padded planes (pixel.cpp:518-541 as called
 * from slicetype.cpp:833-842, 966-976) and weightCostLuma's whole-frame score (:845-858).
 * ------------------------------------------------------------------------------------------ */
template <typename P>
__global__ void __launch_bounds__(256) weight_planes_kernel(const P* __restrict__ src, P* __restrict__ dst, long long n,
                                                            int scale, int round, int shift, int offset, int correction, int maxv)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int val = (short)((int)src[i] << correction);
    const int v = ((scale * val + round) >> shift) + offset;
    dst[i] = (P)min(max(v, 0), maxv);
}

template <typename P>
__global__ void __launch_bounds__(128) weight_cost_kernel(Geom g, const P* __restrict__ fenc0, const P* __restrict__ ref0,
                                                          const int* __restrict__ intraCost, unsigned* result)
{
    __shared__ unsigned s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int cuRaw = blockIdx.x * 16 + grp;
    const bool act = cuRaw < g.ncu;
    const int cu = act ? cuRaw : g.ncu - 1;
    const int X0 = g.mx + 8 * (cu % g.bw), Y0 = g.my + 8 * (cu / g.bw);
    const int satd = groupSatdRows(loadRowAligned(ref0, g.tpr, X0, Y0 + r), loadRowAligned(fenc0, g.tpr, X0, Y0 + r));
    if (r == 0 && act) atomicAdd(&s_sum, (unsigned)min(satd, intraCost[cu]));
    __syncthreads();
    if (threadIdx.x == 0 && s_sum) atomicAdd(result, s_sum);
}

/* ------------------------------------------------------------------------------------------
 * K7/K8: cuTree.  estimateCUPropagateCost (pixel.cpp:931-957) + the MV-directed scatter of
 * Lookahead::estimateCUPropagate (slicetype.cpp:3537-3603); cuTreeFinish (:3784-3796);
 * frameCostRecalculate (:3847-3878).  The reference's CLIP_ADD saturates a uint16 after every
 * add; all addends are >= 0, so min(sum, 65535) taken when the value is read is identical and lets
 * the scatter use plain 32-bit atomics in any order.
 * ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ void cutreePropagateBlock(const Geom& g, int cu, const int* __restrict__ intraCost,
                                                     const unsigned short* __restrict__ lowresCosts,
                                                     const int* __restrict__ invQ, const int* __restrict__ mv0,
                                                     const int* __restrict__ mv1, const int* __restrict__ propagateIn,
                                                     int* ref0, int* ref1, int bipredWeight, double fpsFactor)
{
    const int bw = g.bw, bh = g.bh;
    const int bx = cu % bw, by = cu / bw;
    const double fps = __ddiv_rn(fpsFactor, 256.0);
    const int intra = intraCost[cu];
    const int lc = lowresCosts[cu];
    const int inter = min(intra, lc & LA_LOWRES_COST_MASK);
    const double propagateIntra = (double)(intra * invQ[cu]);
    const int pin = propagateIn ? min(propagateIn[cu], 65535) : 0;
    const double propagateAmount = __dadd_rn((double)pin, __dmul_rn(propagateIntra, fps));
    const double propagateNum = (double)(intra - inter);
    const int amount = (int)__dadd_rn(__ddiv_rn(__dmul_rn(propagateAmount, propagateNum), (double)intra), 0.5);
    if (amount <= 0) return;
    const int lists_used = lc >> LA_LOWRES_COST_SHIFT;
    for (int list = 0; list < 2; list++)
    {
        if (!((lists_used >> list) & 1)) continue;
        int listamount = amount;
        if (lists_used == 3)
            listamount = (listamount * (list ? 64 - bipredWeight : bipredWeight) + 32) >> 6;
        int* ref = list ? ref1 : ref0;
        const MV2 mv = unpackMv(list ? mv1[cu] : mv0[cu]);
        if (!(mv.x | mv.y)) { atomicAdd(ref + cu, listamount); continue; }
        int x = mv.x, y = mv.y;
        const int cux = (x >> 5) + bx, cuy = (y >> 5) + by;
        const int idx0 = cux + cuy * bw;
        x &= 31; y &= 31;
        const int w0 = (32 - y) * (32 - x), w1 = (32 - y) * x, w2 = y * (32 - x), w3 = y * x;
        if (cux < bw && cuy < bh && cux >= 0 && cuy >= 0)             atomicAdd(ref + idx0, (listamount * w0 + 512) >> 10);
        if (cux + 1 < bw && cuy < bh && cux + 1 >= 0 && cuy >= 0)     atomicAdd(ref + idx0 + 1, (listamount * w1 + 512) >> 10);
        if (cux < bw && cuy + 1 < bh && cux >= 0 && cuy + 1 >= 0)     atomicAdd(ref + idx0 + bw, (listamount * w2 + 512) >> 10);
        if (cux + 1 < bw && cuy + 1 < bh && cux + 1 >= 0 && cuy + 1 >= 0) atomicAdd(ref + idx0 + bw + 1, (listamount * w3 + 512) >> 10);
    }
}

__global__ void __launch_bounds__(256) cutree_propagate_kernel(Geom g, const int* __restrict__ intraCost,
                                                               const unsigned short* __restrict__ lowresCosts,
                                                               const int* __restrict__ invQ, const int* __restrict__ mv0,
                                                               const int* __restrict__ mv1, const int* __restrict__ propagateIn,
                                                               int* ref0, int* ref1, int bipredWeight, double fpsFactor)
{
    const int cu = blockIdx.x * blockDim.x + threadIdx.x;
    if (cu >= g.ncu) return;
    cutreePropagateBlock(g, cu, intraCost, lowresCosts, invQ, mv0, mv1, propagateIn, ref0, ref1, bipredWeight, fpsFactor);
}

/* The unreferenced B frames of a mini-GOP only ADD into their two references (integer atomics, any order), so they
 * propagate in one launch (blockIdx.y = frame) instead of one launch each: the cuTree chain is launch-latency bound
 * and the caller waits for it before it can read a frame's qp offsets.  The job table lives in mapped host memory. */
struct CutreeJobDev
{
    const int* intraCost; const unsigned short* lowresCosts; const int* invQ; const int* mv0; const int* mv1;
    int* ref0; int* ref1;
    int* self;          /* the frame's own propagate array: the reference zeroes its first row and uses it as the
                           (empty) incoming amount of an unreferenced frame (slicetype.cpp:3518-3519, 3534-3535) */
    int bipredWeight, pad;
    double fpsFactor;
};

__global__ void __launch_bounds__(256) cutree_propagate_batch_kernel(Geom g, const CutreeJobDev* __restrict__ jobs)
{
    const int cu = blockIdx.x * blockDim.x + threadIdx.x;
    if (cu >= g.ncu) return;
    const CutreeJobDev J = jobs[blockIdx.y];
    if (cu < g.bw) J.self[cu] = 0;      /* observable: that row stays zero in the frame's propagateCost */
    cutreePropagateBlock(g, cu, J.intraCost, J.lowresCosts, J.invQ, J.mv0, J.mv1, NULL, J.ref0, J.ref1, J.bipredWeight, J.fpsFactor);
}

__global__ void __launch_bounds__(256) cutree_finish_kernel(Geom g, const int* __restrict__ intraCost, const int* __restrict__ invQ,
                                                            const int* __restrict__ propagate, const double* __restrict__ qpAq,
                                                            double* __restrict__ qpCuTree, int fpsFactor, double weightdelta,
                                                            double strength)
{
    const int cu = blockIdx.x * blockDim.x + threadIdx.x;
    if (cu >= g.ncu) return;
    if (g.aqBlock == 8)
    {
        /* qg-size 8 (slicetype.cpp:3764-3782): invQ is invQscaleFactor8x8; one ratio for the block's four 8x8 offsets */
        const int intracost = ((intraCost[cu]) / 4 * invQ[cu] + 128) >> 8;
        if (intracost)
        {
            const int propagateCost = (min(propagate[cu], 65535) / 4 * fpsFactor + 128) >> 8;
            const double log2_ratio = __dadd_rn(__dadd_rn(log2((double)(intracost + propagateCost)), -log2((double)intracost)), weightdelta);
            const double d = __dmul_rn(strength, log2_ratio);
            const int fs = 2 * g.bw, i = (cu % g.bw) * 2 + (cu / g.bw) * g.bw * 4;
            qpCuTree[i] = __dadd_rn(qpAq[i], -d);
            qpCuTree[i + 1] = __dadd_rn(qpAq[i + 1], -d);
            qpCuTree[i + fs] = __dadd_rn(qpAq[i + fs], -d);
            qpCuTree[i + fs + 1] = __dadd_rn(qpAq[i + fs + 1], -d);
        }
        return;
    }
    const int intracost = (intraCost[cu] * invQ[cu] + 128) >> 8;
    if (intracost)
    {
        const int propagateCost = (min(propagate[cu], 65535) * fpsFactor + 128) >> 8;
        const double log2_ratio = __dadd_rn(__dadd_rn(log2((double)(intracost + propagateCost)), -log2((double)intracost)), weightdelta);
        qpCuTree[cu] = __dadd_rn(qpAq[cu], -__dmul_rn(strength, log2_ratio));
    }
}

/* the qp offset of lowres block cu: the block's own entry, or with qg-size 8 the mean of its four 8x8 entries
 * (slicetype.cpp:3859-3866, 1414-1421) */
__device__ __forceinline__ double blockQpOffset(const Geom& g, const double* __restrict__ qp, int cu)
{
    if (g.aqBlock != 8) return qp[cu];
    const int fs = 2 * g.bw, i = (cu % g.bw) * 2 + (cu / g.bw) * g.bw * 4;
    return __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(qp[i], qp[i + 1]), qp[i + fs]), qp[i + fs + 1]), 4.0);
}

__global__ void __launch_bounds__(256) cost_recalc_kernel(Geom g, const unsigned short* __restrict__ lowresCosts,
                                                          const double* __restrict__ qpOffset, int* rowSatds,
                                                          unsigned long long* score)
{
    __shared__ unsigned long long s_score;
    if (threadIdx.x == 0) s_score = 0;
    __syncthreads();
    const int cu = blockIdx.x * blockDim.x + threadIdx.x;
    if (cu < g.ncu)
    {
        const int cux = cu % g.bw, cuy = cu / g.bw;
        int cuCost = lowresCosts[cu] & LA_LOWRES_COST_MASK;
        cuCost = (cuCost * exp2fix8(blockQpOffset(g, qpOffset, cu)) + 128) >> 8;
        atomicAdd(&rowSatds[cuy], cuCost);
        if ((cuy > 0 && cuy < g.bh - 1 && cux > 0 && cux < g.bw - 1) || g.bw <= 2 || g.bh <= 2)
            atomicAdd(&s_score, (unsigned long long)cuCost);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_score) atomicAdd(score, s_score);
}

/* The VBV half of Lookahead::getEstimatedPictureCost (slicetype.cpp:1387-1436): the coded estimate's block costs and the
 * intra costs scaled by the block's qp offset, and their sums per CTU row (FrameData::m_rowStat[].satdForVbv /
 * intraSatdForVbv).  The reference rewrites lowresCostForRc / intraCost in place on the host; here the scaled arrays go to
 * scratch that the caller mirrors, the device arrays stay as the lookahead needs them.  qpOffset NULL = no scaling.
 * pirStart/pirEnd: the intra-refresh column range of a P slice (diff term, :1425-1427), -1 = none. */
__global__ void __launch_bounds__(256) vbv_rows_kernel(Geom g, const unsigned short* __restrict__ lowresCosts,
                                                       const int* __restrict__ intraCost, const double* __restrict__ qpOffset,
                                                       int scale, int pirStart, int pirEnd,
                                                       unsigned short* __restrict__ costForRc, int* __restrict__ intraOut,
                                                       unsigned* rowSatd, unsigned* rowIntra)
{
    const int cu = blockIdx.x * blockDim.x + threadIdx.x;
    if (cu >= g.ncu) return;
    const int cuy = cu / g.bw;
    unsigned short c = lowresCosts[cu] & LA_LOWRES_COST_MASK;
    int ic = intraCost[cu];
    if (qpOffset)
    {
        const int f = exp2fix8(blockQpOffset(g, qpOffset, cu));
        c = (unsigned short)((c * f + 128) >> 8);
        ic = (ic * f + 128) >> 8;
    }
    costForRc[cu] = c;
    intraOut[cu] = ic;
    unsigned add = c;
    if (pirStart >= 0)
        add += (unsigned)((pirEnd - pirStart + 1) * (ic - (int)c));
    atomicAdd(&rowSatd[cuy / scale], add);
    atomicAdd(&rowIntra[cuy / scale], (unsigned)ic);
}

/* debug / unit-test kernel: SAD and SATD of n pairs of packed 8x8 blocks (mirrors the reference's
 * check_pixelcmp, source/test/pixelharness.cpp:82).  n must be a multiple of 4 (whole warps). */
template <typename P>
__global__ void __launch_bounds__(128) block_metrics_kernel(const P* __restrict__ a, const P* __restrict__ b, int n,
                                                            int* sadOut, int* satdOut)
{
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int iRaw = blockIdx.x * 16 + grp;
    const int i = min(iRaw, n - 1);
    const Row<P> ra = loadRowPitched(a + (long long)i * 64 + r * 8), rb = loadRowPitched(b + (long long)i * 64 + r * 8);
    int sad = groupSum(sadRow(ra, rb));
    int satd = groupSatdRows(ra, rb);
    /* the same blocks through the 4-lanes-per-block primitives of the search kernel (two rows per lane); a disagreement
     * between the two decompositions is reported as -1, which no oracle value equals */
    __shared__ int s4[2][16];
    {
        const int g4 = (threadIdx.x >> 2) & 15, r2 = (threadIdx.x & 3) * 2;
        const int i4 = min(blockIdx.x * 16 + g4, n - 1);
        const Row<P> fa = loadRowPitched(a + (long long)i4 * 64 + r2 * 8), fb = loadRowPitched(a + (long long)i4 * 64 + r2 * 8 + 8);
        const Row<P> pa = loadRowPitched(b + (long long)i4 * 64 + r2 * 8), pb = loadRowPitched(b + (long long)i4 * 64 + r2 * 8 + 8);
        const int sad4 = group4Sum(sadRows2(fa, fb, pa, pb));
        const int satd4 = group4SatdRows(fa, fb, pa, pb);
        if (threadIdx.x < 64 && (threadIdx.x & 3) == 0) { s4[0][g4] = sad4; s4[1][g4] = satd4; }
    }
    __syncthreads();
    if (s4[0][grp] != sad) sad = -1;
    if (s4[1][grp] != satd) satd = -1;
    if (r == 0 && iRaw < n) { sadOut[i] = sad; satdOut[i] = satd; }
}

/* unit-test kernel for the tiled row fetch + motion compensation: for n (block, mv) pairs returns SATD and
 * SAD of the motion-compensated reference block against the source block, i.e. lowresQPelCost (lowres.h:98-124) */
template <typename P>
__global__ void __launch_bounds__(128) mc_metrics_kernel(Geom g, const P* __restrict__ fenc0, const P* __restrict__ ref0,
                                                         const int* __restrict__ cuIdx, const int* __restrict__ mvs, int n,
                                                         int* sadOut, int* satdOut)
{
    const int grp = threadIdx.x >> 3, r = threadIdx.x & 7;
    const int iRaw = blockIdx.x * 16 + grp;
    const int i = min(iRaw, n - 1);
    const int cu = cuIdx[i];
    const int X0 = g.mx + 8 * (cu % g.bw), Y0 = g.my + 8 * (cu / g.bw);
    const Row<P> fenc = loadRowAligned(fenc0, g.tpr, X0, Y0 + r);
    RefBlock<P> rb = { ref0, g.planeSize, g.tpr, X0, Y0 };
    const Row<P> p = mcRow(rb, mvs[2 * i], mvs[2 * i + 1], r);
    const int sad = groupSum(sadRow(fenc, p));
    const int satd = groupSatdRows(fenc, p);
    if (r == 0 && iRaw < n) { sadOut[i] = sad; satdOut[i] = satd; }
}

} // namespace la
