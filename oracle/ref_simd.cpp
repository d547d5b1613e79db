/* ref_simd.cpp -- TEST / BASELINE INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The image has no nasm, so the reference (DJATOM/x265-aMod 3.6+1) builds with its C primitives only,
 * while BASELINE.json's metric names "x265 asm CPU".  This file gives the reference lookahead an
 * asm-CLASS denominator without its assembly: SSE4.1 intrinsics versions of the primitives the lookahead's
 * hot loop calls, installed into the reference's own (writable, common/primitives.h:436) primitives table:
 *
 *   pu[LUMA_8x8].sad / sad_x3 / sad_x4 / satd        (asm-primitives.cpp:1092-1094,1108,1137-1138)
 *   pu[LUMA_8x8].pixelavg_pp[ALIGNED / NONALIGNED]    (lowresMC's quarter-pel average, lowres.h:71-124)
 *   pu[LUMA_8x8].copy_pp                             (MotionEstimate::setSourcePU, motion.cpp:189)
 *   cu[BLOCK_8x8 / 16x16].var                        (acEnergyCu, slicetype.cpp:60-84)
 *   frameInitLowres                                  (Lowres::init, lowres.cpp:368-370)
 *   propagateCost                                    (estimateCUPropagate, slicetype.cpp:3524-3531)
 *
 * 8x8 blocks are 8 bytes (8-bit) or 16 bytes (16-bit samples) per row, so 128-bit SSE is the natural width
 * (x265's own 8x8 SAD / SATD kernels are SSE2 / SSE4 code too); what these shims lack against the hand-scheduled
 * assembly is tuning, not vector width.  The intra predictors stay C (about 3 % of the lookahead's time).
 * Every shim is verified bit-exact against the C primitive it replaces on the reference's own pixelharness
 * buffer recipe (random / all-min / all-max, test/pixelharness.cpp:31-80) by ref_simd_selftest().
 *
 * Contains no reference source text: the arithmetic definitions are the C primitives' (common/pixel.cpp:40-119,
 * 190-297, 545-557, 605-628, 720-737, 931-957), restated with intrinsics. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <smmintrin.h>

#include "common.h"
#include "primitives.h"
#include "slicetype.h"

using namespace X265_NS;

namespace {

#if HIGH_BIT_DEPTH
/* ---------------------------------------------------------------- 16-bit samples (main10) */
static inline __m128i ldRow(const pixel* p) { return _mm_loadu_si128((const __m128i*)p); }

static inline int hsumW(__m128i accWords)     /* 8 x u16 -> int (each lane <= 8 * 1023) */
{
    __m128i s = _mm_madd_epi16(accWords, _mm_set1_epi16(1));
    s = _mm_add_epi32(s, _mm_srli_si128(s, 8));
    s = _mm_add_epi32(s, _mm_srli_si128(s, 4));
    return _mm_cvtsi128_si32(s);
}

int sad8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    __m128i acc = _mm_setzero_si128();
    for (int y = 0; y < 8; y++, a += sa, b += sb)
        acc = _mm_add_epi16(acc, _mm_abs_epi16(_mm_sub_epi16(ldRow(a), ldRow(b))));
    return hsumW(acc);
}

void sad8x8_x3(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, intptr_t rs, int32_t* res)
{
    __m128i a0 = _mm_setzero_si128(), a1 = a0, a2 = a0;
    for (int y = 0; y < 8; y++, fenc += FENC_STRIDE, r0 += rs, r1 += rs, r2 += rs)
    {
        const __m128i f = ldRow(fenc);
        a0 = _mm_add_epi16(a0, _mm_abs_epi16(_mm_sub_epi16(f, ldRow(r0))));
        a1 = _mm_add_epi16(a1, _mm_abs_epi16(_mm_sub_epi16(f, ldRow(r1))));
        a2 = _mm_add_epi16(a2, _mm_abs_epi16(_mm_sub_epi16(f, ldRow(r2))));
    }
    res[0] = hsumW(a0); res[1] = hsumW(a1); res[2] = hsumW(a2);
}

void sad8x8_x4(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, const pixel* r3, intptr_t rs, int32_t* res)
{
    __m128i a0 = _mm_setzero_si128(), a1 = a0, a2 = a0, a3 = a0;
    for (int y = 0; y < 8; y++, fenc += FENC_STRIDE, r0 += rs, r1 += rs, r2 += rs, r3 += rs)
    {
        const __m128i f = ldRow(fenc);
        a0 = _mm_add_epi16(a0, _mm_abs_epi16(_mm_sub_epi16(f, ldRow(r0))));
        a1 = _mm_add_epi16(a1, _mm_abs_epi16(_mm_sub_epi16(f, ldRow(r1))));
        a2 = _mm_add_epi16(a2, _mm_abs_epi16(_mm_sub_epi16(f, ldRow(r2))));
        a3 = _mm_add_epi16(a3, _mm_abs_epi16(_mm_sub_epi16(f, ldRow(r3))));
    }
    res[0] = hsumW(a0); res[1] = hsumW(a1); res[2] = hsumW(a2); res[3] = hsumW(a3);
}

static inline __m128i diffRow(const pixel* a, const pixel* b) { return _mm_sub_epi16(ldRow(a), ldRow(b)); }

void avg8x8(pixel* dst, intptr_t ds, const pixel* s0, intptr_t ss0, const pixel* s1, intptr_t ss1, int)
{
    for (int y = 0; y < 8; y++, dst += ds, s0 += ss0, s1 += ss1)
        _mm_storeu_si128((__m128i*)dst, _mm_avg_epu16(ldRow(s0), ldRow(s1)));
}

void copy8x8(pixel* dst, intptr_t ds, const pixel* src, intptr_t ss)
{
    for (int y = 0; y < 8; y++, dst += ds, src += ss)
        _mm_storeu_si128((__m128i*)dst, ldRow(src));
}

template <int size>
uint64_t varN(const pixel* p, intptr_t stride)
{
    __m128i sum = _mm_setzero_si128(), sqr = _mm_setzero_si128();
    const __m128i zero = _mm_setzero_si128();
    for (int y = 0; y < size; y++, p += stride)
        for (int x = 0; x < size; x += 8)
        {
            const __m128i v = ldRow(p + x);
            sum = _mm_add_epi32(sum, _mm_add_epi32(_mm_unpacklo_epi16(v, zero), _mm_unpackhi_epi16(v, zero)));
            sqr = _mm_add_epi32(sqr, _mm_madd_epi16(v, v));     /* samples < 2^15: signed multiply is exact */
        }
    sum = _mm_add_epi32(sum, _mm_srli_si128(sum, 8)); sum = _mm_add_epi32(sum, _mm_srli_si128(sum, 4));
    sqr = _mm_add_epi32(sqr, _mm_srli_si128(sqr, 8)); sqr = _mm_add_epi32(sqr, _mm_srli_si128(sqr, 4));
    return (uint32_t)_mm_cvtsi128_si32(sum) + ((uint64_t)(uint32_t)_mm_cvtsi128_si32(sqr) << 32);
}

#define AVG(a, b) _mm_avg_epu16(a, b)
enum { VEC = 8 };       /* output samples per vector step */
static inline void deinterleave(__m128i t0, __m128i t1, __m128i* even, __m128i* odd)
{
    /* t0, t1: 16 consecutive u16; even = elements 0,2,..,14, odd = 1,3,..,15 */
    const __m128i m = _mm_set1_epi32(0x0000ffff);
    *even = _mm_packus_epi32(_mm_and_si128(t0, m), _mm_and_si128(t1, m));
    *odd = _mm_packus_epi32(_mm_srli_epi32(t0, 16), _mm_srli_epi32(t1, 16));
}
#else
/* ---------------------------------------------------------------- 8-bit samples */
static inline __m128i ld2Rows(const pixel* p, intptr_t stride)       /* two 8-byte rows in one register */
{
    return _mm_unpacklo_epi64(_mm_loadl_epi64((const __m128i*)p), _mm_loadl_epi64((const __m128i*)(p + stride)));
}
static inline int hsumSad(__m128i acc) { return _mm_cvtsi128_si32(acc) + _mm_extract_epi16(acc, 4); }

int sad8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    __m128i acc = _mm_setzero_si128();
    for (int y = 0; y < 8; y += 2, a += 2 * sa, b += 2 * sb)
        acc = _mm_add_epi32(acc, _mm_sad_epu8(ld2Rows(a, sa), ld2Rows(b, sb)));
    return hsumSad(acc);
}

void sad8x8_x3(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, intptr_t rs, int32_t* res)
{
    __m128i a0 = _mm_setzero_si128(), a1 = a0, a2 = a0;
    for (int y = 0; y < 8; y += 2, fenc += 2 * FENC_STRIDE, r0 += 2 * rs, r1 += 2 * rs, r2 += 2 * rs)
    {
        const __m128i f = ld2Rows(fenc, FENC_STRIDE);
        a0 = _mm_add_epi32(a0, _mm_sad_epu8(f, ld2Rows(r0, rs)));
        a1 = _mm_add_epi32(a1, _mm_sad_epu8(f, ld2Rows(r1, rs)));
        a2 = _mm_add_epi32(a2, _mm_sad_epu8(f, ld2Rows(r2, rs)));
    }
    res[0] = hsumSad(a0); res[1] = hsumSad(a1); res[2] = hsumSad(a2);
}

void sad8x8_x4(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, const pixel* r3, intptr_t rs, int32_t* res)
{
    __m128i a0 = _mm_setzero_si128(), a1 = a0, a2 = a0, a3 = a0;
    for (int y = 0; y < 8; y += 2, fenc += 2 * FENC_STRIDE, r0 += 2 * rs, r1 += 2 * rs, r2 += 2 * rs, r3 += 2 * rs)
    {
        const __m128i f = ld2Rows(fenc, FENC_STRIDE);
        a0 = _mm_add_epi32(a0, _mm_sad_epu8(f, ld2Rows(r0, rs)));
        a1 = _mm_add_epi32(a1, _mm_sad_epu8(f, ld2Rows(r1, rs)));
        a2 = _mm_add_epi32(a2, _mm_sad_epu8(f, ld2Rows(r2, rs)));
        a3 = _mm_add_epi32(a3, _mm_sad_epu8(f, ld2Rows(r3, rs)));
    }
    res[0] = hsumSad(a0); res[1] = hsumSad(a1); res[2] = hsumSad(a2); res[3] = hsumSad(a3);
}

static inline __m128i diffRow(const pixel* a, const pixel* b)
{
    return _mm_sub_epi16(_mm_cvtepu8_epi16(_mm_loadl_epi64((const __m128i*)a)), _mm_cvtepu8_epi16(_mm_loadl_epi64((const __m128i*)b)));
}

void avg8x8(pixel* dst, intptr_t ds, const pixel* s0, intptr_t ss0, const pixel* s1, intptr_t ss1, int)
{
    for (int y = 0; y < 8; y++, dst += ds, s0 += ss0, s1 += ss1)
        _mm_storel_epi64((__m128i*)dst, _mm_avg_epu8(_mm_loadl_epi64((const __m128i*)s0), _mm_loadl_epi64((const __m128i*)s1)));
}

void copy8x8(pixel* dst, intptr_t ds, const pixel* src, intptr_t ss)
{
    for (int y = 0; y < 8; y++, dst += ds, src += ss)
        _mm_storel_epi64((__m128i*)dst, _mm_loadl_epi64((const __m128i*)src));
}

template <int size>
uint64_t varN(const pixel* p, intptr_t stride)
{
    __m128i sum = _mm_setzero_si128(), sqr = _mm_setzero_si128();
    const __m128i zero = _mm_setzero_si128();
    for (int y = 0; y < size; y++, p += stride)
        for (int x = 0; x < size; x += 8)
        {
            const __m128i v = _mm_loadl_epi64((const __m128i*)(p + x));
            sum = _mm_add_epi32(sum, _mm_sad_epu8(v, zero));
            const __m128i w = _mm_cvtepu8_epi16(v);
            sqr = _mm_add_epi32(sqr, _mm_madd_epi16(w, w));
        }
    sqr = _mm_add_epi32(sqr, _mm_srli_si128(sqr, 8)); sqr = _mm_add_epi32(sqr, _mm_srli_si128(sqr, 4));
    return (uint32_t)_mm_cvtsi128_si32(sum) + ((uint64_t)(uint32_t)_mm_cvtsi128_si32(sqr) << 32);
}

#define AVG(a, b) _mm_avg_epu8(a, b)
enum { VEC = 16 };
static inline void deinterleave(__m128i t0, __m128i t1, __m128i* even, __m128i* odd)
{
    const __m128i m = _mm_set1_epi16(0x00ff);
    *even = _mm_packus_epi16(_mm_and_si128(t0, m), _mm_and_si128(t1, m));
    *odd = _mm_packus_epi16(_mm_srli_epi16(t0, 8), _mm_srli_epi16(t1, 8));
}
#endif

/* ---------------------------------------------------------------- SATD 8x8 = two 8x4 halves, each (sum |H4 D H4|) >> 1
 * over its two 4x4 blocks (pixel.cpp:239-297).  Rows are 8 x int16 differences. */
static inline int satd8x4(__m128i r0, __m128i r1, __m128i r2, __m128i r3)
{
    /* vertical 4-point Hadamard, all 8 columns at once */
    __m128i a0 = _mm_add_epi16(r0, r1), a1 = _mm_sub_epi16(r0, r1), a2 = _mm_add_epi16(r2, r3), a3 = _mm_sub_epi16(r2, r3);
    __m128i b0 = _mm_add_epi16(a0, a2), b1 = _mm_add_epi16(a1, a3), b2 = _mm_sub_epi16(a0, a2), b3 = _mm_sub_epi16(a1, a3);
    /* transpose the two 4x4 blocks (left: words 0-3, right: words 4-7 of each row) */
    __m128i t0 = _mm_unpacklo_epi16(b0, b1), t1 = _mm_unpacklo_epi16(b2, b3);     /* left block, interleaved */
    __m128i t2 = _mm_unpackhi_epi16(b0, b1), t3 = _mm_unpackhi_epi16(b2, b3);     /* right block */
    __m128i c0 = _mm_unpacklo_epi32(t0, t1), c1 = _mm_unpackhi_epi32(t0, t1);     /* left: columns 0,1 | 2,3 */
    __m128i c2 = _mm_unpacklo_epi32(t2, t3), c3 = _mm_unpackhi_epi32(t2, t3);     /* right */
    /* now row j of the transposed left block = 4 words; pair left/right columns in one register */
    __m128i d0 = _mm_unpacklo_epi64(c0, c2), d1 = _mm_unpackhi_epi64(c0, c2), d2 = _mm_unpacklo_epi64(c1, c3), d3 = _mm_unpackhi_epi64(c1, c3);
    /* horizontal Hadamard = vertical one on the transposed data */
    a0 = _mm_add_epi16(d0, d1); a1 = _mm_sub_epi16(d0, d1); a2 = _mm_add_epi16(d2, d3); a3 = _mm_sub_epi16(d2, d3);
    b0 = _mm_abs_epi16(_mm_add_epi16(a0, a2)); b1 = _mm_abs_epi16(_mm_add_epi16(a1, a3));
    b2 = _mm_abs_epi16(_mm_sub_epi16(a0, a2)); b3 = _mm_abs_epi16(_mm_sub_epi16(a1, a3));
    /* coefficients are <= 16 * 1023 < 2^15; widen before summing 32 of them */
    const __m128i one = _mm_set1_epi16(1);
    __m128i s = _mm_add_epi32(_mm_add_epi32(_mm_madd_epi16(b0, one), _mm_madd_epi16(b1, one)),
                              _mm_add_epi32(_mm_madd_epi16(b2, one), _mm_madd_epi16(b3, one)));
    s = _mm_add_epi32(s, _mm_srli_si128(s, 8));
    s = _mm_add_epi32(s, _mm_srli_si128(s, 4));
    return _mm_cvtsi128_si32(s) >> 1;
}

int satd8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    const int top = satd8x4(diffRow(a, b), diffRow(a + sa, b + sb), diffRow(a + 2 * sa, b + 2 * sb), diffRow(a + 3 * sa, b + 3 * sb));
    a += 4 * sa; b += 4 * sb;
    return top + satd8x4(diffRow(a, b), diffRow(a + sa, b + sb), diffRow(a + 2 * sa, b + 2 * sb), diffRow(a + 3 * sa, b + 3 * sb));
}

/* ---------------------------------------------------------------- frame_init_lowres_core (pixel.cpp:605-628)
 * FILTER(a,b,c,d) = avg(avg(a,b), avg(c,d)) with avg(x,y) = (x+y+1)>>1 = pavgb / pavgw. */
void lowresSimd(const pixel* src0, pixel* dst0, pixel* dsth, pixel* dstv, pixel* dstc, intptr_t srcStride, intptr_t dstStride,
                int width, int height)
{
    for (int y = 0; y < height; y++)
    {
        const pixel* s0 = src0;
        const pixel* s1 = s0 + srcStride;
        const pixel* s2 = s1 + srcStride;
        int x = 0;
        for (; x + VEC <= width; x += VEC)
        {
            __m128i T[2][2];        /* [top / bottom pair][first / second half] of the horizontally averaged rows */
            for (int h = 0; h < 2; h++)
            {
                const int o = 2 * x + h * VEC;
                const __m128i r0 = _mm_loadu_si128((const __m128i*)(s0 + o)), r0n = _mm_loadu_si128((const __m128i*)(s0 + o + 1));
                const __m128i r1 = _mm_loadu_si128((const __m128i*)(s1 + o)), r1n = _mm_loadu_si128((const __m128i*)(s1 + o + 1));
                const __m128i r2 = _mm_loadu_si128((const __m128i*)(s2 + o)), r2n = _mm_loadu_si128((const __m128i*)(s2 + o + 1));
                T[0][h] = AVG(AVG(r0, r1), AVG(r0n, r1n));
                T[1][h] = AVG(AVG(r1, r2), AVG(r1n, r2n));
            }
            __m128i e, o;
            deinterleave(T[0][0], T[0][1], &e, &o);
            _mm_storeu_si128((__m128i*)(dst0 + x), e); _mm_storeu_si128((__m128i*)(dsth + x), o);
            deinterleave(T[1][0], T[1][1], &e, &o);
            _mm_storeu_si128((__m128i*)(dstv + x), e); _mm_storeu_si128((__m128i*)(dstc + x), o);
        }
        for (; x < width; x++)
        {
#define FILTER(a, b, c, d) ((((a + b + 1) >> 1) + ((c + d + 1) >> 1) + 1) >> 1)
            dst0[x] = FILTER(s0[2 * x], s1[2 * x], s0[2 * x + 1], s1[2 * x + 1]);
            dsth[x] = FILTER(s0[2 * x + 1], s1[2 * x + 1], s0[2 * x + 2], s1[2 * x + 2]);
            dstv[x] = FILTER(s1[2 * x], s2[2 * x], s1[2 * x + 1], s2[2 * x + 1]);
            dstc[x] = FILTER(s1[2 * x + 1], s2[2 * x + 1], s1[2 * x + 2], s2[2 * x + 2]);
#undef FILTER
        }
        src0 += srcStride * 2;
        dst0 += dstStride; dsth += dstStride; dstv += dstStride; dstc += dstStride;
    }
}

/* ---------------------------------------------------------------- estimateCUPropagateCost (pixel.cpp:931-957), two
 * blocks per step in double precision; the same operations in the same order as the scalar code */
void propagateSimd(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                   const int32_t* invQscales, const double* fpsFactor, int len)
{
    const double fps = *fpsFactor / 256;
    const __m128d vfps = _mm_set1_pd(fps), half = _mm_set1_pd(0.5);
    int i = 0;
    for (; i + 2 <= len; i += 2)
    {
        const __m128i intra = _mm_loadl_epi64((const __m128i*)(intraCosts + i));
        const __m128i inter = _mm_and_si128(_mm_set_epi32(0, 0, interCosts[i + 1], interCosts[i]), _mm_set1_epi32(LOWRES_COST_MASK));
        const __m128i interMin = _mm_min_epi32(intra, inter);
        const __m128i prod = _mm_mullo_epi32(intra, _mm_loadl_epi64((const __m128i*)(invQscales + i)));
        const __m128d propagateIntra = _mm_cvtepi32_pd(prod);
        const __m128d pin = _mm_cvtepi32_pd(_mm_set_epi32(0, 0, propagateIn[i + 1], propagateIn[i]));
        const __m128d amount = _mm_add_pd(pin, _mm_mul_pd(propagateIntra, vfps));
        const __m128d num = _mm_cvtepi32_pd(_mm_sub_epi32(intra, interMin));
        const __m128d den = _mm_cvtepi32_pd(intra);
        const __m128d r = _mm_add_pd(_mm_div_pd(_mm_mul_pd(amount, num), den), half);
        const __m128i out = _mm_cvttpd_epi32(r);
        _mm_storel_epi64((__m128i*)(dst + i), out);
    }
    for (; i < len; i++)
    {
        const int intraCost = intraCosts[i];
        const int interCost = X265_MIN(intraCosts[i], interCosts[i] & LOWRES_COST_MASK);
        const double propagateIntra = intraCost * invQscales[i];
        const double propagateAmount = (double)propagateIn[i] + propagateIntra * fps;
        const double propagateNum = (double)(intraCost - interCost);
        dst[i] = (int)(propagateAmount * propagateNum / (double)intraCost + 0.5);
    }
}

/* ---------------------------------------------------------------- install / verify */
struct Saved
{
    bool valid, installed;
    pixelcmp_t sad, satd; pixelcmp_x3_t sad_x3; pixelcmp_x4_t sad_x4;
    pixelavg_pp_t avg[2]; copy_pp_t copy; var_t var8, var16; downscale_t lowres; cutree_propagate_cost propagate;
} g_saved;

void saveC()
{
    if (g_saved.valid) return;
    EncoderPrimitives& p = primitives;
    g_saved.sad = p.pu[LUMA_8x8].sad; g_saved.satd = p.pu[LUMA_8x8].satd; g_saved.sad_x3 = p.pu[LUMA_8x8].sad_x3;
    g_saved.sad_x4 = p.pu[LUMA_8x8].sad_x4; g_saved.avg[0] = p.pu[LUMA_8x8].pixelavg_pp[0]; g_saved.avg[1] = p.pu[LUMA_8x8].pixelavg_pp[1];
    g_saved.copy = p.pu[LUMA_8x8].copy_pp; g_saved.var8 = p.cu[BLOCK_8x8].var; g_saved.var16 = p.cu[BLOCK_16x16].var;
    g_saved.lowres = p.frameInitLowres; g_saved.propagate = p.propagateCost;
    g_saved.valid = true;
}

uint32_t rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return s >> 8; }

} // namespace

extern "C" {

/* 1 when the running CPU can execute the shims */
int ref_simd_supported(void) { return __builtin_cpu_supports("sse4.1") ? 1 : 0; }

/* on != 0: install the intrinsics shims into the reference's primitives table; 0: put the C primitives back.
 * Call after the encoder was opened (x265_setup_primitives fills the table).  Returns 1 if the shims are active. */
int ref_simd_install(int on)
{
    saveC();
    EncoderPrimitives& p = primitives;
    if (on && ref_simd_supported())
    {
        p.pu[LUMA_8x8].sad = sad8x8; p.pu[LUMA_8x8].satd = satd8x8; p.pu[LUMA_8x8].sad_x3 = sad8x8_x3; p.pu[LUMA_8x8].sad_x4 = sad8x8_x4;
        p.pu[LUMA_8x8].pixelavg_pp[0] = avg8x8; p.pu[LUMA_8x8].pixelavg_pp[1] = avg8x8; p.pu[LUMA_8x8].copy_pp = copy8x8;
        p.cu[BLOCK_8x8].var = varN<8>; p.cu[BLOCK_16x16].var = varN<16>;
        p.frameInitLowres = lowresSimd; p.propagateCost = propagateSimd;
        g_saved.installed = true;
    }
    else
    {
        p.pu[LUMA_8x8].sad = g_saved.sad; p.pu[LUMA_8x8].satd = g_saved.satd; p.pu[LUMA_8x8].sad_x3 = g_saved.sad_x3;
        p.pu[LUMA_8x8].sad_x4 = g_saved.sad_x4; p.pu[LUMA_8x8].pixelavg_pp[0] = g_saved.avg[0]; p.pu[LUMA_8x8].pixelavg_pp[1] = g_saved.avg[1];
        p.pu[LUMA_8x8].copy_pp = g_saved.copy; p.cu[BLOCK_8x8].var = g_saved.var8; p.cu[BLOCK_16x16].var = g_saved.var16;
        p.frameInitLowres = g_saved.lowres; p.propagateCost = g_saved.propagate;
        g_saved.installed = false;
    }
    return g_saved.installed ? 1 : 0;
}

/* Differential test of every shim against the C primitive it replaces, on the pixelharness recipe: random buffers,
 * all-minimum and all-maximum buffers, random strides and offsets (test/pixelharness.cpp:31-80, 82-173).
 * Returns the number of mismatches (0 = bit-exact). */
int ref_simd_selftest(int iterations)
{
    saveC();
    if (!ref_simd_supported()) return -1;
    const int W = 160, H = 80, N = W * H;
    const int maxv = (1 << X265_DEPTH) - 1;
    pixel* buf[3][2];
    for (int k = 0; k < 3; k++)
        for (int j = 0; j < 2; j++)
            buf[k][j] = (pixel*)malloc(N * sizeof(pixel) + 64);
    uint32_t seed = 12345;
    for (int i = 0; i < N; i++)
    {
        buf[0][0][i] = (pixel)(rnd(seed) & maxv); buf[0][1][i] = (pixel)(rnd(seed) & maxv);
        buf[1][0][i] = 0; buf[1][1][i] = (pixel)maxv;
        buf[2][0][i] = (pixel)maxv; buf[2][1][i] = 0;
    }
    int bad = 0;
    ALIGN_VAR_32(pixel, fenc[FENC_STRIDE * 8]);
    for (int it = 0; it < iterations; it++)
    {
        const int k = it % 3;
        const pixel* a = buf[k][0]; const pixel* b = buf[k][1];
        const int sa = 8 + (int)(rnd(seed) % 100), sb = 8 + (int)(rnd(seed) % 100);
        const int oa = (int)(rnd(seed) % (N - 16 * 110)), ob = (int)(rnd(seed) % (N - 8 * 110 - 16));    /* var16 reads 16 rows */
        bad += g_saved.sad(a + oa, sa, b + ob, sb) != sad8x8(a + oa, sa, b + ob, sb);
        bad += g_saved.satd(a + oa, sa, b + ob, sb) != satd8x8(a + oa, sa, b + ob, sb);
        for (int y = 0; y < 8; y++) memcpy(fenc + y * FENC_STRIDE, a + oa + y * sa, 8 * sizeof(pixel));
        int32_t r0[4], r1[4];
        g_saved.sad_x3(fenc, b + ob, b + ob + 1, b + ob + 2 + sb, sb, r0); sad8x8_x3(fenc, b + ob, b + ob + 1, b + ob + 2 + sb, sb, r1);
        bad += memcmp(r0, r1, 12) != 0;
        g_saved.sad_x4(fenc, b + ob, b + ob + 1, b + ob + 7, b + ob + sb, sb, r0); sad8x8_x4(fenc, b + ob, b + ob + 1, b + ob + 7, b + ob + sb, sb, r1);
        bad += memcmp(r0, r1, 16) != 0;
        pixel d0[64], d1[64];
        g_saved.avg[0](d0, 8, a + oa, sa, b + ob, sb, 32); avg8x8(d1, 8, a + oa, sa, b + ob, sb, 32);
        bad += memcmp(d0, d1, sizeof(d0)) != 0;
        g_saved.copy(d0, 8, a + oa, sa); copy8x8(d1, 8, a + oa, sa);
        bad += memcmp(d0, d1, sizeof(d0)) != 0;
        bad += g_saved.var8(a + oa, sa) != varN<8>(a + oa, sa);
        if (sa >= 16) bad += g_saved.var16(a + oa, sa) != varN<16>(a + oa, sa);
    }
    /* lowres downscale on whole buffers, several widths incl. a ragged tail */
    for (int k = 0; k < 3; k++)
        for (int w = 8; w <= 72; w += (k == 0 ? 8 : 40))
        {
            const int h = 30;
            pixel* o[2][4];
            for (int j = 0; j < 2; j++) for (int q = 0; q < 4; q++) o[j][q] = (pixel*)calloc(80 * h + 32, sizeof(pixel));
            g_saved.lowres(buf[k][k == 0 ? 0 : 1], o[0][0], o[0][1], o[0][2], o[0][3], W, 80, w, h);
            lowresSimd(buf[k][k == 0 ? 0 : 1], o[1][0], o[1][1], o[1][2], o[1][3], W, 80, w, h);
            for (int q = 0; q < 4; q++) bad += memcmp(o[0][q], o[1][q], 80 * h * sizeof(pixel)) != 0;
            for (int j = 0; j < 2; j++) for (int q = 0; q < 4; q++) free(o[j][q]);
        }
    /* cuTree propagate: random costs, both parities of the length */
    for (int it = 0; it < 200; it++)
    {
        const int len = 1 + (int)(rnd(seed) % 61);
        int d0[64], d1[64]; uint16_t pin[64], inter[64]; int32_t intra[64], invq[64];
        for (int i = 0; i < len; i++)
        {
            pin[i] = (uint16_t)rnd(seed); inter[i] = (uint16_t)rnd(seed);
            intra[i] = 1 + (int)(rnd(seed) % 40000); invq[i] = 1 + (int)(rnd(seed) % 2000);
        }
        const double fps = 0.01 + (rnd(seed) % 1000) / 1010.0;
        g_saved.propagate(d0, pin, intra, inter, invq, &fps, len); propagateSimd(d1, pin, intra, inter, invq, &fps, len);
        bad += memcmp(d0, d1, len * sizeof(int)) != 0;
    }
    for (int k = 0; k < 3; k++) for (int j = 0; j < 2; j++) free(buf[k][j]);
    return bad;
}

} // extern "C"
