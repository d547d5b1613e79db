/* la_oracle.c -- CPU restatement of the reference lookahead's block-level arithmetic.
 * TEST INFRASTRUCTURE ONLY; see la_oracle.h for scope, pinning status and usage rules.
 * Citations are file:line under /root/reference/source/. */
#include "la_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define OR_MAXV ((1 << OR_DEPTH) - 1)
#define OR_COST_MAX (1 << 28)            /* encoder/motion.h:65 */
#define OR_LOWRES_COST_MASK 16383        /* encoder/slicetype.h:41 */
#define OR_LOWRES_COST_SHIFT 14

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

int or_depth(void) { return OR_DEPTH; }

/* common/lowres.cpp:72-97, common/picyuv.cpp:87-88 */
void or_geom_init(or_geom* g, int picW, int picH, int maxCUSize)
{
    g->picW = picW; g->picH = picH;
    int lw = picW / 2, lh = picH / 2;
    g->mx = maxCUSize + 32; g->my = maxCUSize + 16;
    g->stride = lw + 2 * g->mx;
    if (g->stride & 31) g->stride += 32 - (g->stride & 31);
    g->bw = (lw + 7) >> 3; g->bh = (lh + 7) >> 3; g->ncu = g->bw * g->bh;
    g->w = g->bw * 8; g->h = g->bh * 8;
    g->planeLines = g->h + 2 * g->my;
    g->planeSize = (int64_t)g->stride * g->planeLines;
    g->padOffset = (int64_t)g->stride * g->my + g->mx;
    g->rowsPerSlice = 0; g->qg8 = 0;
    /* --hme, level 0: the 1/16-resolution planes live in buffers of half the size, addressed with half the stride and
     * extended by half the margins (common/lowres.cpp:165-183, 378-388); the level-0 block grid comes from the SOURCE
     * size (encoder/slicetype.cpp:998-999) */
    g->w4 = g->w / 2; g->h4 = g->h / 2;
    g->bw4 = ((picW / 4) + 7) >> 3; g->bh4 = ((picH / 4) + 7) >> 3;
    g->mx4 = g->mx / 2; g->my4 = g->my / 2;
    g->stride4 = g->stride / 2;
    g->planeSize4 = g->planeSize / 2;
    g->padOffset4 = g->padOffset / 2;
}

/* common/constants.cpp:34-90: lambda = 2^(qp/6 - 2) * 2^(depth-8); X265_LOOKAHEAD_QP = 12 + 6*(depth-8)
 * (common/common.h:209-213) -> 1, 16, 256 for 8/10/12 bit */
int or_lookahead_lambda(void) { return 1 << (2 * (OR_DEPTH - 8)); }

/* encoder/bitcost.cpp:46-54 (cost row) and :98-113 (bit sizes: the C `log` of a float argument, i.e. double arithmetic up to
 * the store into the float array) */
void or_build_mvcost(uint16_t* table, int half)
{
    double lambda = (double)or_lookahead_lambda();
    float log2_2 = (float)(2.0f / log((double)2.0f));
    for (int i = 0; i <= half; i++)
    {
        float argument = (float)(i + 1);
        double value = log((double)argument) * log2_2 + 1.718f;
        float bits = i ? (float)value : 0.718f;
        double c = bits * lambda + 0.5f;
        if (c > (double)((1 << 15) - 1)) c = (double)((1 << 15) - 1);
        table[half + i] = table[half - i] = (uint16_t)c;
    }
}

/* common/constants.cpp:552-558 */
static const uint8_t exp2_lut[64] = {
    0, 3, 6, 8, 11, 14, 17, 20, 23, 26, 29, 32, 36, 39, 42, 45,
    48, 52, 55, 58, 62, 65, 69, 72, 76, 80, 83, 87, 91, 94, 98, 102,
    106, 110, 114, 118, 122, 126, 130, 135, 139, 143, 147, 152, 156, 161, 165, 170,
    175, 179, 184, 189, 194, 198, 203, 208, 214, 219, 224, 229, 234, 240, 245, 250 };

/* common/common.cpp:96-103 */
int or_exp2fix8(double x)
{
    int i = (int)(x * (-64.f / 6.f) + 512.5f);
    if (i < 0) return 0;
    if (i > 1023) return 0xffff;
    return (exp2_lut[i & 63] + 256) << (i >> 6) >> 8;
}

/* common/pixel.cpp:40-56 */
int or_sad8x8(const or_pixel* a, int sa, const or_pixel* b, int sb)
{
    int sum = 0;
    for (int y = 0; y < 8; y++, a += sa, b += sb)
        for (int x = 0; x < 8; x++)
            sum += abs((int)a[x] - (int)b[x]);
    return sum;
}

/* SATD of an 8x8 block as the reference defines it (pixel.cpp:239-297): two 8x4 halves, each
 * the sum of |4x4 Hadamard coefficients| of its two 4x4 blocks, halved (>>1) per 8x4. */
static int hadamard4x4_abs(const or_pixel* a, int sa, const or_pixel* b, int sb)
{
    int d[4][4], t[4][4], sum = 0;
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            d[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    for (int y = 0; y < 4; y++)
    {
        int s01 = d[y][0] + d[y][1], d01 = d[y][0] - d[y][1];
        int s23 = d[y][2] + d[y][3], d23 = d[y][2] - d[y][3];
        t[y][0] = s01 + s23; t[y][1] = s01 - s23; t[y][2] = d01 + d23; t[y][3] = d01 - d23;
    }
    for (int x = 0; x < 4; x++)
    {
        int s01 = t[0][x] + t[1][x], d01 = t[0][x] - t[1][x];
        int s23 = t[2][x] + t[3][x], d23 = t[2][x] - t[3][x];
        sum += abs(s01 + s23) + abs(s01 - s23) + abs(d01 + d23) + abs(d01 - d23);
    }
    return sum;
}

int or_satd8x8(const or_pixel* a, int sa, const or_pixel* b, int sb)
{
    int total = 0;
    for (int row = 0; row < 8; row += 4)
    {
        int s = hadamard4x4_abs(a + row * sa, sa, b + row * sb, sb) +
                hadamard4x4_abs(a + row * sa + 4, sa, b + row * sb + 4, sb);
        total += s >> 1;
    }
    return total;
}

/* ------------------------------------------------------------------ lowres planes */

static inline int src_at(const or_pixel* src, int stride, int W, int H, int x, int y)
{
    if (x > W - 1) x = W - 1;
    if (y > H - 1) y = H - 1;
    return src[(int64_t)y * stride + x];
}

#define OR_FILTER(a, b, c, d) ((((a + b + 1) >> 1) + ((c + d + 1) >> 1) + 1) >> 1)

/* common/pixel.cpp:605-628 + common/lowres.cpp:367-376 + pixel.cpp:1044-1058 + ipfilter.cpp:59-77 */
void or_lowres_init(const or_geom* g, const or_pixel* srcY, int srcStride, or_pixel* buf)
{
    or_pixel* pl[4];
    for (int i = 0; i < 4; i++) pl[i] = buf + i * g->planeSize + g->padOffset;
    const int W = g->picW, H = g->picH, st = g->stride;
    for (int y = 0; y < g->h; y++)
        for (int x = 0; x < g->w; x++)
        {
            int a00 = src_at(srcY, srcStride, W, H, 2 * x, 2 * y),     a01 = src_at(srcY, srcStride, W, H, 2 * x + 1, 2 * y),     a02 = src_at(srcY, srcStride, W, H, 2 * x + 2, 2 * y);
            int a10 = src_at(srcY, srcStride, W, H, 2 * x, 2 * y + 1), a11 = src_at(srcY, srcStride, W, H, 2 * x + 1, 2 * y + 1), a12 = src_at(srcY, srcStride, W, H, 2 * x + 2, 2 * y + 1);
            int a20 = src_at(srcY, srcStride, W, H, 2 * x, 2 * y + 2), a21 = src_at(srcY, srcStride, W, H, 2 * x + 1, 2 * y + 2), a22 = src_at(srcY, srcStride, W, H, 2 * x + 2, 2 * y + 2);
            pl[0][y * st + x] = (or_pixel)OR_FILTER(a00, a10, a01, a11);
            pl[1][y * st + x] = (or_pixel)OR_FILTER(a01, a11, a02, a12);
            pl[2][y * st + x] = (or_pixel)OR_FILTER(a10, a20, a11, a21);
            pl[3][y * st + x] = (or_pixel)OR_FILTER(a11, a21, a12, a22);
        }
    for (int i = 0; i < 4; i++)
    {
        or_pixel* p = pl[i];
        for (int y = 0; y < g->h; y++)
            for (int x = 0; x < g->mx; x++)
            {
                p[y * st - g->mx + x] = p[y * st];
                p[y * st + g->w + x] = p[y * st + g->w - 1];
            }
        /* rows above/below copy one whole buffer row of `stride` pixels */
        or_pixel* top = p - g->mx;
        for (int y = 0; y < g->my; y++)
            memcpy(top - (y + 1) * st, top, st * sizeof(or_pixel));
        or_pixel* bot = p - g->mx + (g->h - 1) * st;
        for (int y = 0; y < g->my; y++)
            memcpy(bot + (y + 1) * st, bot, st * sizeof(or_pixel));
    }
}

/* --hme: Lowres::init's second call of the downscale primitive, lowresPlane[0] -> the four 1/16-resolution planes, and their
 * border extension by half the margins (common/lowres.cpp:378-388, pixel.cpp:605-628).  `plane0` is lowresPlane[0] (margins
 * already extended: the filter reads one column / row beyond the plane); `buf4` is 4 * planeSize4 pixels, zeroed by the caller. */
void or_lowerres_init(const or_geom* g, const or_pixel* plane0, or_pixel* buf4)
{
    or_pixel* pl[4];
    for (int i = 0; i < 4; i++) pl[i] = buf4 + i * g->planeSize4 + g->padOffset4;
    const int st = g->stride, st4 = g->stride4;
    for (int y = 0; y < g->h4; y++)
    {
        const or_pixel* s0 = plane0 + (int64_t)2 * y * st;
        const or_pixel* s1 = s0 + st;
        const or_pixel* s2 = s1 + st;
        for (int x = 0; x < g->w4; x++)
        {
            pl[0][y * st4 + x] = (or_pixel)OR_FILTER(s0[2 * x], s1[2 * x], s0[2 * x + 1], s1[2 * x + 1]);
            pl[1][y * st4 + x] = (or_pixel)OR_FILTER(s0[2 * x + 1], s1[2 * x + 1], s0[2 * x + 2], s1[2 * x + 2]);
            pl[2][y * st4 + x] = (or_pixel)OR_FILTER(s1[2 * x], s2[2 * x], s1[2 * x + 1], s2[2 * x + 1]);
            pl[3][y * st4 + x] = (or_pixel)OR_FILTER(s1[2 * x + 1], s2[2 * x + 1], s1[2 * x + 2], s2[2 * x + 2]);
        }
    }
    for (int i = 0; i < 4; i++)
    {
        or_pixel* p = pl[i];
        for (int y = 0; y < g->h4; y++)
            for (int x = 0; x < g->mx4; x++)
            {
                p[y * st4 - g->mx4 + x] = p[y * st4];
                p[y * st4 + g->w4 + x] = p[y * st4 + g->w4 - 1];
            }
        /* extendPicBorder copies (width + 2 * marginX) pixels per row above / below (pixel.cpp:1044-1058 via picyuv) */
        const int rowLen = g->w4 + 2 * g->mx4;
        or_pixel* top = p - g->mx4;
        for (int y = 0; y < g->my4; y++)
            memcpy(top - (y + 1) * st4, top, rowLen * sizeof(or_pixel));
        or_pixel* bot = p - g->mx4 + (g->h4 - 1) * st4;
        for (int y = 0; y < g->my4; y++)
            memcpy(bot + (y + 1) * st4, bot, rowLen * sizeof(or_pixel));
    }
}

/* ------------------------------------------------------------------ adaptive quant */

/* pixel_var<N> + acEnergyVar (pixel.cpp:720-737, slicetype.cpp:49-57); source is addressed
 * with replicate clamping (blocks may overhang the picture into PicYuv's padding) */
static uint32_t block_energy(const or_pixel* p, int stride, int W, int H, int bx, int by, int size, int shift,
                             uint64_t* wpSum, uint64_t* wpSsd)
{
    uint32_t sum = 0, sqr = 0;
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++)
        {
            uint32_t v = (uint32_t)src_at(p, stride, W, H, bx + x, by + y);
            sum += v; sqr += v * v;
        }
    *wpSum += sum; *wpSsd += sqr;
    return sqr - (uint32_t)(((uint64_t)sum * sum) >> shift);
}

/* aq-mode 4 / 5: edgeFilter + computeEdge (encoder/slicetype.cpp:98-223) over the luma plane.  edge / theta are W x H (the
 * reference's buffers are zero outside the picture; callers treat anything outside as 0).  Gaussian 5x5 (integer, / 159) inside a
 * 2-sample border, the border keeps the source; Sobel-like gradients on that inside a 1-sample border: edge = white (the
 * largest sample value) where sqrtf(gH^2 + gV^2) >= EDGE_THRESHOLD else 0, the border keeps the SOURCE sample; theta = the
 * gradient angle in whole degrees 0..180, 0 on the border. */
static void edge_filter(const or_pixel* y, int strideY, int W, int H, or_pixel* edge, or_pixel* theta)
{
    or_pixel* gauss = (or_pixel*)malloc(sizeof(or_pixel) * (size_t)W * H);
    for (int r = 0; r < H; r++)
        for (int c = 0; c < W; c++)
        {
            const or_pixel* s = y + (int64_t)r * strideY + c;
            int v = s[0];
            if (r >= 2 && c >= 2 && r < H - 2 && c < W - 2)
            {
                static const int k[5][5] = { {2, 4, 5, 4, 2}, {4, 9, 12, 9, 4}, {5, 12, 15, 12, 5}, {4, 9, 12, 9, 4}, {2, 4, 5, 4, 2} };
                int acc = 0;
                for (int j = -2; j <= 2; j++)
                    for (int i = -2; i <= 2; i++)
                        acc += k[j + 2][i + 2] * s[(int64_t)j * strideY + i];
                v = (or_pixel)(acc / 159);
            }
            gauss[(size_t)r * W + c] = (or_pixel)v;
            edge[(size_t)r * W + c] = s[0];
            theta[(size_t)r * W + c] = 0;
        }
    /* EDGE_THRESHOLD (slicetype.h:64-69): 255 for 8-bit, 1023 for EVERY high bit depth; it is also the white sample */
    const int white = OR_DEPTH > 8 ? 1023 : 255;
    const float threshold = (float)white;
    for (int r = 1; r < H - 1; r++)
        for (int c = 1; c < W - 1; c++)
        {
            const or_pixel* p = gauss + (size_t)r * W + c;
            const float gH = (float)(-3 * p[-W - 1] + 3 * p[-W + 1] - 10 * p[-1] + 10 * p[1] - 3 * p[W - 1] + 3 * p[W + 1]);
            const float gV = (float)(-3 * p[-W - 1] - 10 * p[-W] - 3 * p[-W + 1] + 3 * p[W - 1] + 10 * p[W] + 3 * p[W + 1]);
            const float mag = sqrtf(gH * gH + gV * gV);
            const float radians = (float)atan2(gV, gH);
            float th = (float)((radians * 180) / 3.14159265);        /* PI, slicetype.h:70 */
            if (th < 0) th = 180 + th;
            theta[(size_t)r * W + c] = (or_pixel)th;
            edge[(size_t)r * W + c] = (or_pixel)(mag >= threshold ? white : 0);
        }
    free(gauss);
}

/* LookaheadTLD::edgeDensityCu (slicetype.cpp:236-258): variance of the edge image's block and the block's mean angle; like
 * every acEnergyVar call it adds the block's sum / sum of squares to the weightp statistics of plane 0 (:54-55) */
static uint32_t edge_density(const or_pixel* edge, const or_pixel* theta, int W, int H, int bx, int by, int size, int shift,
                             uint32_t* avgAngle, uint64_t* wpSum, uint64_t* wpSsd)
{
    uint32_t sum = 0, sqr = 0, ang = 0;
    for (int y = 0; y < size; y++)
        for (int x = 0; x < size; x++)
        {
            if (bx + x >= W || by + y >= H) continue;       /* the reference's buffers are zeroed outside the picture (:170-172) */
            const uint32_t v = edge[(size_t)(by + y) * W + bx + x];
            sum += v; sqr += v * v;
            ang += theta[(size_t)(by + y) * W + bx + x];
        }
    *avgAngle = ang / (uint32_t)(size * size);
    *wpSum += sum; *wpSsd += sqr;
    return sqr - (uint32_t)(((uint64_t)sum * sum) >> shift);
}

/* encoder/slicetype.cpp:452-713 */
void or_aq_frame(const or_geom* g, const or_pixel* y, int strideY, const or_pixel* u, const or_pixel* v,
                 int strideC, int aqMode, double aqStrength, int bWeightP,
                 double* qpAqOffset, double* qpCuTreeOffset, int32_t* invQscaleFactor,
                 uint32_t* blockEnergy, uint64_t wp_ssd[3], uint64_t wp_sum[3])
{
    const int W = g->picW, H = g->picH;
    /* slicetype.cpp:459-472: qg-size 8 works on 8x8 luma / 4x4 chroma blocks; the arrays hold blockCount entries but the
     * loops below visit ceil(W / incr) * ceil(H / incr) blocks with a running index */
    const int qg8 = g->qg8;
    const int blockCount = qg8 ? 4 * g->ncu : g->ncu;
    const int incr = qg8 ? 8 : 16;
    const float modeOneConst = qg8 ? 11.427f : 14.427f, modeTwoConst = qg8 ? 8.f : 11.f;
    for (int i = 0; i < 3; i++) wp_ssd[i] = wp_sum[i] = 0;
    const int visited = ((W + incr - 1) / incr) * ((H + incr - 1) / incr);
    uint32_t* energy = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(visited > blockCount ? visited : blockCount));
    const int edgeAq = (aqMode == 4 || aqMode == 5) && aqStrength != 0;
    or_pixel *edge = NULL, *theta = NULL;
    uint8_t* inclined = NULL;
    if (edgeAq)
    {
        edge = (or_pixel*)malloc(sizeof(or_pixel) * (size_t)W * H); theta = (or_pixel*)malloc(sizeof(or_pixel) * (size_t)W * H);
        inclined = (uint8_t*)calloc((size_t)(visited > blockCount ? visited : blockCount), 1);
        edge_filter(y, strideY, W, H, edge, theta);
    }
    int n = 0;
    for (int by = 0; by < H; by += incr)
        for (int bx = 0; bx < W; bx += incr, n++)
        {
            uint32_t e = block_energy(y, strideY, W, H, bx, by, incr, qg8 ? 6 : 8, &wp_sum[0], &wp_ssd[0]);
            if (u && v)
            {
                e += block_energy(u, strideC, (W + 1) >> 1, (H + 1) >> 1, bx >> 1, by >> 1, incr / 2, qg8 ? 4 : 6, &wp_sum[1], &wp_ssd[1]);
                e += block_energy(v, strideC, (W + 1) >> 1, (H + 1) >> 1, bx >> 1, by >> 1, incr / 2, qg8 ? 4 : 6, &wp_sum[2], &wp_ssd[2]);
            }
            if (blockEnergy) blockEnergy[n] = e;
            if (edgeAq)
            {
                /* :568-585: a block with edges takes its edge density instead of its energy, and is marked when the mean
                 * gradient angle lies within 15 degrees of a diagonal (EDGE_INCLINATION 45) */
                uint32_t avgAngle = 0;
                const uint32_t density = edge_density(edge, theta, W, H, bx, by, incr, qg8 ? 6 : 8, &avgAngle, &wp_sum[0], &wp_ssd[0]);
                if (density)
                {
                    e = density;
                    inclined[n] = (avgAngle >= 30 && avgAngle <= 60) || (avgAngle >= 120 && avgAngle <= 150);
                }
            }
            energy[n] = e;
        }

    if (aqMode == 0 || aqStrength == 0)
    {
        if (aqMode && aqStrength == 0)
            for (int i = 0; i < blockCount; i++) { qpAqOffset[i] = qpCuTreeOffset[i] = 0; invQscaleFactor[i] = 256; }
    }
    else
    {
        double avg_adj_pow2 = 0, avg_adj = 0, qp_adj = 0, bias_strength = 0, strength = 0;
        if (aqMode == 2 || aqMode == 3 || aqMode == 4 || aqMode == 5)
        {
            double bit_depth_correction = 1.f / (1 << (2 * (OR_DEPTH - 8)));
            for (int i = 0; i < visited; i++)
            {
                qp_adj = pow(energy[i] * bit_depth_correction + 1, 0.1);
                qpCuTreeOffset[i] = qp_adj;
                avg_adj += qp_adj;
                avg_adj_pow2 += qp_adj * qp_adj;
            }
            avg_adj /= blockCount;
            avg_adj_pow2 /= blockCount;
            strength = aqStrength * avg_adj;
            avg_adj = avg_adj - 0.5f * (avg_adj_pow2 - modeTwoConst) / avg_adj;
            bias_strength = 1.0 /* aqBiasStrength default */ * aqStrength;
        }
        else
            strength = aqStrength * 1.0397f;
        for (int i = 0; i < visited; i++)
        {
            if (aqMode == 3)
            {
                qp_adj = qpCuTreeOffset[i];
                qp_adj = strength * (qp_adj - avg_adj) + bias_strength * (1.f - modeTwoConst / (qp_adj * qp_adj));
            }
            else if (aqMode == 2)
            {
                qp_adj = qpCuTreeOffset[i];
                qp_adj = strength * (qp_adj - avg_adj);
            }
            else if (aqMode == 4)
            {
                qp_adj = qpCuTreeOffset[i];
                if (inclined[i] && (qp_adj - avg_adj > 0))
                    qp_adj = ((strength + 0.5 /* AQ_EDGE_BIAS */) * (qp_adj - avg_adj));
                else
                    qp_adj = strength * (qp_adj - avg_adj);
            }
            else if (aqMode == 5)
            {
                qp_adj = qpCuTreeOffset[i];
                double dark_bias = bias_strength * (1.f - modeTwoConst / (qp_adj * qp_adj)) / 10.f;
                if (inclined[i] && (qp_adj - avg_adj > 0))
                    qp_adj = ((strength + 0.5) * (qp_adj - avg_adj));
                else
                    qp_adj = strength * (qp_adj - avg_adj);
                qp_adj += dark_bias;
            }
            else
            {
                uint32_t e = energy[i] > 1 ? energy[i] : 1;
                qp_adj = strength * (log2((double)e) - (modeOneConst + 2 * (OR_DEPTH - 8)));
            }
            qpAqOffset[i] = qp_adj;
            qpCuTreeOffset[i] = qp_adj;
            invQscaleFactor[i] = or_exp2fix8(qp_adj);
        }
    }
    if (bWeightP)
    {
        int maxCol = ((W + 8) >> 4) << 4, maxRow = ((H + 8) >> 4) << 4;
        int wd[3] = { maxCol, maxCol >> 1, maxCol >> 1 }, ht[3] = { maxRow, maxRow >> 1, maxRow >> 1 };
        for (int i = 0; i < 3; i++)
        {
            uint64_t sum = wp_sum[i], ssd = wp_ssd[i];
            wp_ssd[i] = ssd - (sum * sum + (uint64_t)(wd[i] * ht[i]) / 2) / (uint64_t)(wd[i] * ht[i]);
        }
    }
    free(energy); free(edge); free(theta); free(inclined);
}

/* --fades: the tail of calcAdaptiveQuantFrame (encoder/slicetype.cpp:697-712).  A second acEnergyCu pass over the picture
 * fills blockVariance and sums it into frameVariance: the row sum is never reset between rows, its division by maxCol is an
 * integer one, and maxCol / maxRow are the values the weightp block above left behind (the size rounded to 16, :683-684) when
 * weightp is on.  acEnergyVar's side effect (:54-55) adds every block's sum / ssd to wp_sum / wp_ssd once more, after they
 * were finalised.  Call after or_aq_frame. */
double or_fade_variance(const or_geom* g, const or_pixel* y, int strideY, const or_pixel* u, const or_pixel* v, int strideC,
                        int bWeightP, uint64_t wp_ssd[3], uint64_t wp_sum[3])
{
    const int W = g->picW, H = g->picH;
    const int qg8 = g->qg8, incr = qg8 ? 8 : 16;
    const int maxCol = bWeightP ? ((W + 8) >> 4) << 4 : W, maxRow = bWeightP ? ((H + 8) >> 4) << 4 : H;
    uint64_t rowVariance = 0;
    double frameVariance = 0;
    for (int by = 0; by < maxRow; by += incr)
    {
        for (int bx = 0; bx < maxCol; bx += incr)
        {
            uint32_t e = block_energy(y, strideY, W, H, bx, by, incr, qg8 ? 6 : 8, &wp_sum[0], &wp_ssd[0]);
            if (u && v)
            {
                e += block_energy(u, strideC, (W + 1) >> 1, (H + 1) >> 1, bx >> 1, by >> 1, incr / 2, qg8 ? 4 : 6, &wp_sum[1], &wp_ssd[1]);
                e += block_energy(v, strideC, (W + 1) >> 1, (H + 1) >> 1, bx >> 1, by >> 1, incr / 2, qg8 ? 4 : 6, &wp_sum[2], &wp_ssd[2]);
            }
            rowVariance += e;
        }
        frameVariance += (double)(rowVariance / (uint64_t)maxCol);
    }
    return frameVariance / maxRow;
}

/* ------------------------------------------------------------------ --hist-scenecut picture statistics */

/* pixel_var<N> + acEnergyVarHist (pixel.cpp:720-737, slicetype.cpp:90-96) on a block inside the (replicate-padded) picture */
static uint32_t hist_block_var(const or_pixel* p, int stride, int W, int H, int bx, int by, int size, int shift)
{
    uint32_t sum = 0, sqr = 0;
    for (int yy = 0; yy < size; yy++)
        for (int xx = 0; xx < size; xx++)
        {
            uint32_t v = (uint32_t)src_at(p, stride, W, H, bx + xx, by + yy);
            sum += v; sqr += v * v;
        }
    return sqr - (uint32_t)(((uint64_t)sum * sum) >> shift);
}

/* LookaheadTLD::calculateHistogram (slicetype.cpp:1549-1571) */
static uint64_t hist_accumulate(const or_pixel* src, uint32_t width, uint32_t height, int stride, int dsFactor, uint32_t* histogram)
{
    uint64_t sum = 0;
    for (uint32_t vy = 0; vy < height; vy += dsFactor)
    {
        for (uint32_t hx = 0; hx < width; hx += dsFactor)
        {
            ++histogram[src[hx] & 255];     /* 8-bit samples; the mask only keeps a (refused) high-bit-depth build in bounds */
            sum += src[hx];
        }
        src += (stride << (dsFactor >> 1));
    }
    return sum;
}

void or_hist_stats(const or_geom* g, const or_pixel* y, int strideY, const or_pixel* u, const or_pixel* v, int strideC,
                   const or_pixel* plane0, or_hist_stats_t* out)
{
    const int W = g->picW, H = g->picH;
    memset(out, 0, sizeof(*out));
    /* quarter-sampled luma: frame_lowres_core on lowresPlane[0] (lowres.cpp:35-51, 392-402) */
    const int qW = W / 4, qH = H / 4;
    or_pixel* q = (or_pixel*)malloc(sizeof(or_pixel) * (size_t)(qW > 0 ? qW : 1) * (size_t)(qH > 0 ? qH : 1));
    for (int yy = 0; yy < qH; yy++)
        for (int xx = 0; xx < qW; xx++)
        {
            const or_pixel* s0 = plane0 + (int64_t)(2 * yy) * g->stride, *s1 = s0 + g->stride;
            q[yy * qW + xx] = (or_pixel)OR_FILTER(s0[2 * xx], s1[2 * xx], s0[2 * xx + 1], s1[2 * xx + 1]);
        }
    uint64_t sumLuma = 0, sumCb = 0, sumCr = 0;
    {   /* computeIntensityHistogramBinsLuma (:1649-1694) */
        const uint32_t segW = (uint32_t)qW / 4, segH = (uint32_t)qH / 4;
        for (uint32_t wi = 0; wi < 4; wi++)
            for (uint32_t hi = 0; hi < 4; hi++)
            {
                uint32_t* hist = out->histogram[wi][hi][0];
                for (int b = 0; b < 256; b++) hist[b] = 1;
                const uint32_t offW = wi == 3 ? (uint32_t)qW - 4 * segW : 0, offH = hi == 3 ? (uint32_t)qH - 4 * segH : 0;
                const uint64_t sum = hist_accumulate(q + wi * segW + (int64_t)(hi * segH) * qW, segW + offW, segH + offH, qW, 1, hist);
                /* (the rounding term multiplies two WIDTHS in the reference, :1683) */
                out->avgIntensitySeg[wi][hi][0] = (uint8_t)((sum + (((segW + offW) * (segW + offH)) >> 1)) / ((segW + offW) * (segH + offH)));
                sumLuma += sum << 4;
                for (int b = 0; b < 256; b++) hist[b] <<= 4;
            }
    }
    {   /* computeIntensityHistogramBinsChroma (:1576-1644) */
        const uint32_t segW = (uint32_t)W / 4, segH = (uint32_t)H / 4;
        const int dsFactor = 4;
        for (uint32_t wi = 0; wi < 4; wi++)
            for (uint32_t hi = 0; hi < 4; hi++)
            {
                const uint32_t offW = wi == 3 ? (uint32_t)W - 4 * segW : 0, offH = hi == 3 ? (uint32_t)H - 4 * segH : 0;
                for (int pl = 1; pl <= 2; pl++)
                {
                    uint32_t* hist = out->histogram[wi][hi][pl];
                    for (int b = 0; b < 256; b++) hist[b] = 1;
                    const or_pixel* src = (pl == 1 ? u : v) + ((wi * segW) >> 1) + (int64_t)((hi * segH) >> 1) * strideC;
                    uint64_t sum = hist_accumulate(src, (segW + offW) >> 1, (segH + offH) >> 1, strideC, dsFactor, hist);
                    sum <<= dsFactor;
                    if (pl == 1) sumCb += sum; else sumCr += sum;
                    /* (the V divisor adds the HEIGHT offset to the width in the reference, :1632) */
                    const uint32_t den = pl == 1 ? ((segW + offW) * (segH + offH)) >> 2 : ((segW + offH) * (segH + offH)) >> 2;
                    out->avgIntensitySeg[wi][hi][pl] = (uint8_t)((sum + (((segW + offW) * (segH + offH)) >> 3)) / den);
                    for (int b = 0; b < 256; b++) hist[b] <<= dsFactor;
                }
            }
    }
    /* collectPictureStatistics (:1698-1724) */
    out->avgIntensity[0] = (uint8_t)((sumLuma + (((uint64_t)W * H) >> 1)) / ((uint64_t)W * H));
    out->avgIntensity[1] = (uint8_t)((sumCb + (((uint64_t)W * H) >> 3)) / (((uint64_t)W * H) >> 2));
    out->avgIntensity[2] = (uint8_t)((sumCr + (((uint64_t)W * H) >> 3)) / (((uint64_t)W * H) >> 2));
    {   /* computePictureStatistics (:1457-1544): per block row the variances are summed, divided by the width and CUT to 16 bits */
        uint64_t tot = 0;
        for (int by = 0; by < H; by += 8)
        {
            uint64_t row = 0;
            for (int bx = 0; bx < W; bx += 8) row += hist_block_var(y, strideY, W, H, bx, by, 8, 6);
            tot += (uint16_t)(row / (uint64_t)W);
        }
        out->picAvgVariance[0] = (uint16_t)(tot / (uint64_t)H);
        const int cW = W >> 1, cH = H >> 1;
        for (int pl = 1; pl <= 2; pl++)
        {
            tot = 0;
            for (int by = 0; by < cH; by += 4)
            {
                uint64_t row = 0;
                for (int bx = 0; bx < cW; bx += 4) row += hist_block_var(pl == 1 ? u : v, strideC, (W + 1) >> 1, (H + 1) >> 1, bx, by, 4, 4);
                tot += (uint16_t)(row / (uint64_t)cW);
            }
            out->picAvgVariance[pl] = (uint16_t)(tot / (uint64_t)cH);
        }
    }
    free(q);
}

/* ------------------------------------------------------------------ intra */

/* common/constants.cpp:561-567 */
static const uint8_t intraFilterFlags[35] = {
    0x38, 0x00,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30, 0x20, 0x00, 0x20, 0x30, 0x30, 0x30, 0x30, 0x30, 0x30,
    0x38 };

/* common/intrapred.cpp:31-51 (tuSize 8): nb[0]=top-left, [1..16]=top, [17..32]=left */
static void intra_filter8(const or_pixel* s, or_pixel* f)
{
    int topLeft = s[0], topLast = s[16], leftLast = s[32];
    for (int i = 1; i < 16; i++) f[i] = (or_pixel)(((s[i] << 1) + s[i - 1] + s[i + 1] + 2) >> 2);
    f[16] = (or_pixel)topLast;
    f[0] = (or_pixel)(((topLeft << 1) + s[1] + s[17] + 2) >> 2);
    f[17] = (or_pixel)(((s[17] << 1) + topLeft + s[18] + 2) >> 2);
    for (int i = 18; i < 32; i++) f[i] = (or_pixel)(((s[i] << 1) + s[i - 1] + s[i + 1] + 2) >> 2);
    f[32] = (or_pixel)leftLast;
}

/* common/intrapred.cpp:53-85 */
static void pred_dc8(or_pixel* dst, const or_pixel* s)
{
    int dc = 8;
    for (int i = 0; i < 8; i++) dc += s[1 + i] + s[17 + i];
    dc /= 16;
    for (int i = 0; i < 64; i++) dst[i] = (or_pixel)dc;
    const or_pixel* above = s + 1; const or_pixel* left = s + 17;
    dst[0] = (or_pixel)((above[0] + left[0] + 2 * dst[0] + 2) >> 2);
    for (int x = 1; x < 8; x++) dst[x] = (or_pixel)((above[x] + 3 * dst[x] + 2) >> 2);
    for (int y = 1; y < 8; y++) dst[y * 8] = (or_pixel)((left[y] + 3 * dst[y * 8] + 2) >> 2);
}

/* common/intrapred.cpp:87-100 */
static void pred_planar8(or_pixel* dst, const or_pixel* s)
{
    const or_pixel* above = s + 1; const or_pixel* left = s + 17;
    int topRight = above[8], bottomLeft = left[8];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            dst[y * 8 + x] = (or_pixel)(((7 - x) * left[y] + (7 - y) * above[x] + (x + 1) * topRight + (y + 1) * bottomLeft + 8) >> 4);
}

/* common/intrapred.cpp:102-204 (width 8, bFilter = 1) */
static void pred_ang8(or_pixel* dst, const or_pixel* s0, int mode)
{
    static const int8_t angleTable[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
    static const int16_t invAngleTable[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };
    int hor = mode < 18;
    or_pixel nb[33];
    const or_pixel* s = s0;
    if (hor)
    {
        nb[0] = s0[0];
        for (int i = 0; i < 16; i++) { nb[1 + i] = s0[17 + i]; nb[17 + i] = s0[1 + i]; }
        s = nb;
    }
    int angleOffset = hor ? 10 - mode : mode - 26;
    int angle = angleTable[8 + angleOffset];
    if (!angle)
    {
        for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) dst[y * 8 + x] = s[1 + x];
        int topLeft = s[0], top = s[1];
        for (int y = 0; y < 8; y++)
        {
            int v = (int16_t)(top + ((s[17 + y] - topLeft) >> 1));
            dst[y * 8] = (or_pixel)(v < 0 ? 0 : (v > OR_MAXV ? OR_MAXV : v));
        }
    }
    else
    {
        or_pixel refBuf[64]; const or_pixel* ref;
        if (angle < 0)
        {
            int nbProjected = -((8 * angle) >> 5) - 1;
            or_pixel* rp = refBuf + nbProjected + 1;
            int invAngle = invAngleTable[-angleOffset - 1], invAngleSum = 128;
            for (int i = 0; i < nbProjected; i++) { invAngleSum += invAngle; rp[-2 - i] = s[16 + (invAngleSum >> 8)]; }
            for (int i = 0; i < 9; i++) rp[-1 + i] = s[i];
            ref = rp;
        }
        else
            ref = s + 1;
        int angleSum = 0;
        for (int y = 0; y < 8; y++)
        {
            angleSum += angle;
            int off = angleSum >> 5, frac = angleSum & 31;
            for (int x = 0; x < 8; x++)
                dst[y * 8 + x] = frac ? (or_pixel)(((32 - frac) * ref[off + x] + frac * ref[off + x + 1] + 16) >> 5) : ref[off + x];
        }
    }
    if (hor)
        for (int y = 0; y < 7; y++)
            for (int x = y + 1; x < 8; x++)
            { or_pixel t = dst[y * 8 + x]; dst[y * 8 + x] = dst[x * 8 + y]; dst[x * 8 + y] = t; }
}

/* encoder/slicetype.cpp:715-824 */
void or_intra_estimate(const or_geom* g, const or_pixel* plane0, const int32_t* invQ,
                       int32_t* intraCost, uint8_t* intraMode, uint16_t* lowresCosts00,
                       int32_t* rowSatds00, int64_t* costEst00, int64_t* costEstAq00)
{
    const int lambda = or_lookahead_lambda();
    const int intraPenalty = 5 * lambda, lowresPenalty = 4;
    const int st = g->stride;
    int64_t costEst = 0, costEstAq = 0;
    or_pixel pred[64], fenc[64], nbA[33], nbF[33];
    for (int cuY = 0; cuY < g->bh; cuY++)
    {
        rowSatds00[cuY] = 0;
        for (int cuX = 0; cuX < g->bw; cuX++)
        {
            const int cuXY = cuX + cuY * g->bw;
            const or_pixel* pix = plane0 + 8 * cuX + (int64_t)8 * cuY * st;
            for (int y = 0; y < 8; y++) memcpy(fenc + 8 * y, pix + y * st, 8 * sizeof(or_pixel));
            const or_pixel* c = pix - st - 1;
            memcpy(nbA, c, 17 * sizeof(or_pixel));
            for (int i = 1; i <= 16; i++) nbA[16 + i] = c[(int64_t)i * st];
            intra_filter8(nbA, nbF);

            int cost, icost = OR_COST_MAX, ilow = 0;
            pred_dc8(pred, nbA);
            cost = or_satd8x8(fenc, 8, pred, 8);
            if (cost < icost) { icost = cost; ilow = 1; }
            pred_planar8(pred, nbF);
            cost = or_satd8x8(fenc, 8, pred, 8);
            if (cost < icost) { icost = cost; ilow = 0; }
            int acost = OR_COST_MAX, alow = 4;
            for (int mode = 5; mode < 35; mode += 5)
            {
                pred_ang8(pred, (intraFilterFlags[mode] & 8) ? nbF : nbA, mode);
                cost = or_satd8x8(fenc, 8, pred, 8);
                if (cost < acost) { acost = cost; alow = mode; }
            }
            for (int dist = 2; dist >= 1; dist--)
            {
                int minus = alow - dist, plus = alow + dist;
                pred_ang8(pred, (intraFilterFlags[minus] & 8) ? nbF : nbA, minus);
                cost = or_satd8x8(fenc, 8, pred, 8);
                if (cost < acost) { acost = cost; alow = minus; }
                pred_ang8(pred, (intraFilterFlags[plus] & 8) ? nbF : nbA, plus);
                cost = or_satd8x8(fenc, 8, pred, 8);
                if (cost < acost) { acost = cost; alow = plus; }
            }
            if (acost < icost) { icost = acost; ilow = alow; }
            icost += intraPenalty + lowresPenalty;
            lowresCosts00[cuXY] = (uint16_t)imin(icost, OR_LOWRES_COST_MASK);
            intraCost[cuXY] = icost;
            intraMode[cuXY] = (uint8_t)ilow;
            int scored = (cuX > 0 && cuX < g->bw - 1 && cuY > 0 && cuY < g->bh - 1) || g->bw <= 2 || g->bh <= 2;
            int icostAq = (scored && invQ) ? ((icost * invQ[cuXY] + 128) >> 8) : icost;
            if (scored) { costEst += icost; costEstAq += icostAq; }
            rowSatds00[cuY] += icostAq;
        }
    }
    *costEst00 = costEst;
    *costEstAq00 = costEstAq;
}

/* ------------------------------------------------------------------ motion search */

typedef struct { int x, y; } or_mv;

typedef struct
{
    const or_geom* g;
    or_pixel fenc[64];
    const or_pixel* ref[4];      /* lowresPlane[0..3] + block offset */
    int stride;
    const uint16_t* mvcost;      /* centre */
    or_mv mvp;
} or_me;

static inline int mvcost_q(const or_me* m, int qx, int qy)   /* bitcost.h:46 */
{
    return (uint16_t)(m->mvcost[qx - m->mvp.x] + m->mvcost[qy - m->mvp.y]);
}

/* ReferencePlanes::lowresMC (common/lowres.h:71-96): returns pointer+stride of the 8x8 prediction */
static const or_pixel* lowres_mc(const or_me* m, int qx, int qy, or_pixel* buf, int* outStride)
{
    const int st = m->stride;
    if ((qx | qy) & 1)
    {
        int hpelA = (qy & 2) | ((qx & 2) >> 1);
        const or_pixel* a = m->ref[hpelA] + (qx >> 2) + (int64_t)(qy >> 2) * st;
        int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        int hpelB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const or_pixel* b = m->ref[hpelB] + (qx2 >> 2) + (int64_t)(qy2 >> 2) * st;
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                buf[y * 8 + x] = (or_pixel)((a[y * st + x] + b[y * st + x] + 1) >> 1);
        *outStride = 8;
        return buf;
    }
    int hpel = (qy & 2) | ((qx & 2) >> 1);
    *outStride = st;
    return m->ref[hpel] + (qx >> 2) + (int64_t)(qy >> 2) * st;
}

/* ReferencePlanes::lowresQPelCost (common/lowres.h:98-124) */
static int qpel_cost(const or_me* m, int qx, int qy, int useSatd)
{
    or_pixel buf[64]; int st;
    const or_pixel* p = lowres_mc(m, qx, qy, buf, &st);
    return useSatd ? or_satd8x8(m->fenc, 8, p, st) : or_sad8x8(m->fenc, 8, p, st);
}

static inline int sad_fpel(const or_me* m, int x, int y)
{
    return or_sad8x8(m->fenc, 8, m->ref[0] + x + (int64_t)y * m->stride, m->stride);
}

static const or_mv hex2[8] = { {-1, -2}, {-2, 0}, {-1, 2}, {1, 2}, {2, 0}, {1, -2}, {-1, -2}, {-2, 0} };   /* motion.cpp:64 */
static const uint8_t mod6m1[8] = { 5, 0, 1, 2, 3, 4, 5, 0 };                                               /* motion.cpp:65 */
static const or_mv square1[9] = { {0, 0}, {0, -1}, {0, 1}, {-1, 0}, {1, 0}, {-1, -1}, {-1, 1}, {1, -1}, {1, 1} }; /* :66 */

static const or_mv hex4[16] = { {0, -4}, {0, 4}, {-2, -3}, {2, -3}, {-4, -2}, {4, -2}, {-4, -1}, {4, -1},
                                 {-4, 0}, {4, 0}, {-4, 1}, {4, 1}, {-4, 2}, {4, 2}, {-2, 3}, {2, 3} };                    /* motion.cpp:67-73 */

#define OR_DIA_SEARCH 0
#define OR_HEX_SEARCH 1
#define OR_UMH_SEARCH 2
#define OR_STAR_SEARCH 3
#define OR_FULL_SEARCH 5

/* offsets for the two-point search around a distance-1 result (motion.cpp:74-84) */
static const or_mv star_offsets[16] = { {-1, 0}, {0, -1}, {-1, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {-1, -1},
                                        {1, -1}, {1, 1}, {-1, 0}, {0, 1}, {-1, 1}, {1, 1}, {1, 0}, {0, 1} };

static inline int in_range(or_mv v, or_mv lo, or_mv hi) { return v.x >= lo.x && v.x <= hi.x && v.y >= lo.y && v.y <= hi.y; }   /* mv.h:104 */

/* MotionEstimate::StarPatternSearch (motion.cpp:387-630): a growing star around the entry point (which stays the centre);
 * every point either measured unconditionally (the whole star lies inside the range) or behind the reference's own,
 * partial, range tests */
typedef struct { int bcost; or_mv bmv; int point, dist; } or_star;

static void star_pt(or_me* m, or_star* s, int mx, int my, int point, int dist)
{
    int cost = sad_fpel(m, mx, my) + mvcost_q(m, mx << 2, my << 2);
    if (cost < s->bcost) { s->bcost = cost; s->bmv.x = mx; s->bmv.y = my; s->point = point; s->dist = dist; }
}

static void star_pattern(or_me* m, or_mv mvmin, or_mv mvmax, or_star* s, int earlyExitIters, int merange)
{
    const or_mv o = s->bmv;
    int saved = s->bcost, rounds = 0;
    for (int dist = 1; dist <= 8; dist <<= 1)
    {
        const int top = o.y - dist, bottom = o.y + dist, left = o.x - dist, right = o.x + dist;
        const int top2 = o.y - (dist >> 1), bottom2 = o.y + (dist >> 1), left2 = o.x - (dist >> 1), right2 = o.x + (dist >> 1);
        const int in = top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y;
        saved = s->bcost;
        if (dist == 1)
        {
            if (in || top >= mvmin.y) star_pt(m, s, o.x, top, 2, dist);
            if (in || left >= mvmin.x) star_pt(m, s, left, o.y, 4, dist);
            if (in || right <= mvmax.x) star_pt(m, s, right, o.y, 5, dist);
            if (in || bottom <= mvmax.y) star_pt(m, s, o.x, bottom, 7, dist);
        }
        else
        {
            if (in || top >= mvmin.y) star_pt(m, s, o.x, top, 2, dist);
            if (in || top2 >= mvmin.y)
            {
                if (in || left2 >= mvmin.x) star_pt(m, s, left2, top2, 1, dist >> 1);
                if (in || right2 <= mvmax.x) star_pt(m, s, right2, top2, 3, dist >> 1);
            }
            if (in || left >= mvmin.x) star_pt(m, s, left, o.y, 4, dist);
            if (in || right <= mvmax.x) star_pt(m, s, right, o.y, 5, dist);
            if (in || bottom2 <= mvmax.y)
            {
                if (in || left2 >= mvmin.x) star_pt(m, s, left2, bottom2, 6, dist >> 1);
                if (in || right2 <= mvmax.x) star_pt(m, s, right2, bottom2, 8, dist >> 1);
            }
            if (in || bottom <= mvmax.y) star_pt(m, s, o.x, bottom, 7, dist);
        }
        if (s->bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
    for (int dist = 16; dist <= (int16_t)merange; dist <<= 1)
    {
        const int top = o.y - dist, bottom = o.y + dist, left = o.x - dist, right = o.x + dist;
        const int in = top >= mvmin.y && left >= mvmin.x && right <= mvmax.x && bottom <= mvmax.y;
        saved = s->bcost;
        if (in || top >= mvmin.y) star_pt(m, s, o.x, top, 0, dist);
        if (in || left >= mvmin.x) star_pt(m, s, left, o.y, 0, dist);
        if (in || right <= mvmax.x) star_pt(m, s, right, o.y, 0, dist);
        if (in || bottom <= mvmax.y) star_pt(m, s, o.x, bottom, 0, dist);
        for (int index = 1; index < 4; index++)
        {
            const int yT = top + (dist >> 2) * index, yB = bottom - (dist >> 2) * index;
            const int xL = o.x - (dist >> 2) * index, xR = o.x + (dist >> 2) * index;
            if (in || yT >= mvmin.y)
            {
                if (in || xL >= mvmin.x) star_pt(m, s, xL, yT, 0, dist);
                if (in || xR <= mvmax.x) star_pt(m, s, xR, yT, 0, dist);
            }
            if (in || yB <= mvmax.y)
            {
                if (in || xL >= mvmin.x) star_pt(m, s, xL, yB, 0, dist);
                if (in || xR <= mvmax.x) star_pt(m, s, xR, yB, 0, dist);
            }
        }
        if (s->bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
}

static void star_two_points(or_me* m, or_mv mvmin, or_mv mvmax, or_star* s)
{
    const or_mv c = s->bmv;
    for (int k = 0; k < 2; k++)
    {
        or_mv v = { c.x + star_offsets[(s->point - 1) * 2 + k].x, c.y + star_offsets[(s->point - 1) * 2 + k].y };
        if (in_range(v, mvmin, mvmax))
        {
            int cost = sad_fpel(m, v.x, v.y) + mvcost_q(m, v.x << 2, v.y << 2);
            if (cost < s->bcost) { s->bcost = cost; s->bmv = v; }
        }
    }
}

/* X265_STAR_SEARCH (motion.cpp:1157-1264) */
static void star_search(or_me* m, or_mv mvmin, or_mv mvmax, int merange, int* bcost, or_mv* bmv)
{
    or_star s = { *bcost, *bmv, 0, 0 };
    star_pattern(m, mvmin, mvmax, &s, 3, merange);
    int stop = 0;
    if (s.dist == 1)
    {
        if (!s.point) stop = 1;
        else
        {
            int saved = s.bcost;
            star_two_points(m, mvmin, mvmax, &s);
            if (s.bcost == saved) stop = 1;
        }
    }
    if (!stop)
    {
        if (s.dist > 5)     /* RasterDistance */
            for (int y = mvmin.y; y <= mvmax.y; y += 5)
                for (int x = mvmin.x; x <= mvmax.x; x += 5)
                {
                    if (x + 15 <= mvmax.x)
                    {
                        for (int k = 0; k < 4; k++)
                        {
                            /* the fourth of a group is charged mvcost(tmv << 3) (:1219, sic) */
                            int sh = k == 3 ? 3 : 2, xk = x + 5 * k;
                            int cost = sad_fpel(m, xk, y) + mvcost_q(m, xk << sh, y << sh);
                            if (cost < s.bcost) { s.bcost = cost; s.bmv.x = xk; s.bmv.y = y; }
                        }
                        x += 15;
                    }
                    else
                    {
                        int cost = sad_fpel(m, x, y) + mvcost_q(m, x << 2, y << 2);
                        if (cost < s.bcost) { s.bcost = cost; s.bmv.x = x; s.bmv.y = y; }
                    }
                }
        while (s.dist > 0)
        {
            s.dist = 0; s.point = 0;
            star_pattern(m, mvmin, mvmax, &s, 32, merange);
            if (s.dist == 1)
            {
                if (s.point) star_two_points(m, mvmin, mvmax, &s);
                break;
            }
        }
    }
    *bcost = s.bcost; *bmv = s.bmv;
}


/* MotionEstimate::motionEstimate, lowres path (numCandidates 0, subpelRefine 1): the integer search `method` (DIA / HEX / UMH,
 * encoder/motion.cpp:842-1160) over `merange`, then the lowres subpel refinement (:1473-1528).  Without --hme the lookahead
 * always runs HEX over 16 (slicetype.cpp:4096,4170). */
static int motion_estimate_ex(or_me* m, or_mv mvmin, or_mv mvmax, or_mv qmvp, int merange, int method, or_mv* out)
{
    m->mvp = qmvp;
    or_mv qmin = { mvmin.x << 2, mvmin.y << 2 }, qmax = { mvmax.x << 2, mvmax.y << 2 };
    or_mv pmv = { imax(imin(qmvp.x, qmax.x), qmin.x), imax(imin(qmvp.y, qmax.y), qmin.y) };
    or_mv bestpre = pmv;
    int bprecost = qpel_cost(m, pmv.x, pmv.y, 0);
    or_mv bmv = { (pmv.x + 2) >> 2, (pmv.y + 2) >> 2 };
    int bcost = bprecost;
    if ((pmv.x & 3) | (pmv.y & 3))
        bcost = sad_fpel(m, bmv.x, bmv.y) + mvcost_q(m, bmv.x << 2, bmv.y << 2);
    if (pmv.x | pmv.y)
    {
        int cost = sad_fpel(m, 0, 0) + mvcost_q(m, 0, 0);
        if (cost < bcost)
        {
            bcost = cost;
            bmv.x = 0;
            bmv.y = imax(imin(0, mvmax.y), mvmin.y);
        }
    }
    pmv.x = (pmv.x + 2) >> 2; pmv.y = (pmv.y + 2) >> 2;      /* motion.cpp:839 */
    int hexRefine = method == OR_HEX_SEARCH;
#define YOK(dy) ((bmv.y + (dy) >= mvmin.y) & (bmv.y + (dy) <= mvmax.y))
#define COSTAT(dx, dy) (sad_fpel(m, bmv.x + (dx), bmv.y + (dy)) + mvcost_q(m, (bmv.x + (dx)) << 2, (bmv.y + (dy)) << 2))
    if (method == OR_DIA_SEARCH)
    {   /* diamond, radius 1 (motion.cpp:845-868) */
        bcost <<= 4;
        int i = merange;
        do
        {
            int c0 = COSTAT(0, -1), c1 = COSTAT(0, 1), c2 = COSTAT(-1, 0), c3 = COSTAT(1, 0);
            if (YOK(-1)) { if ((c0 << 4) + 1 < bcost) bcost = (c0 << 4) + 1; }
            if (YOK(1))  { if ((c1 << 4) + 3 < bcost) bcost = (c1 << 4) + 3; }
            if ((c2 << 4) + 4 < bcost) bcost = (c2 << 4) + 4;
            if ((c3 << 4) + 12 < bcost) bcost = (c3 << 4) + 12;
            if (!(bcost & 15))
                break;
            bmv.x -= (int32_t)((uint32_t)bcost << 28) >> 30;
            bmv.y -= (int32_t)((uint32_t)bcost << 30) >> 30;
            bcost &= ~15;
        }
        while (--i && in_range(bmv, mvmin, mvmax));
        bcost >>= 4;
    }
    else if (method == OR_UMH_SEARCH)
    {   /* uneven multi-hexagon (motion.cpp:971-1160), numCandidates == 0 */
        or_mv omv = bmv;
        int ucost1, ucost2, cross_start = 1;
        /* COST_MV / COST_MV_X4 (motion.cpp:263-269, 309-330): the latter measures around omv and range-checks y only */
#define COST_MV(mx, my) { int cost_ = sad_fpel(m, (mx), (my)) + mvcost_q(m, (mx) << 2, (my) << 2); \
                          if (cost_ < bcost) { bcost = cost_; bmv.x = (mx); bmv.y = (my); } }
#define COST_MV_1(dx, dy) { int cost_ = sad_fpel(m, omv.x + (dx), omv.y + (dy)) + mvcost_q(m, (omv.x + (dx)) << 2, (omv.y + (dy)) << 2); \
                            if ((omv.y + (dy) >= mvmin.y) & (omv.y + (dy) <= mvmax.y)) \
                                if (cost_ < bcost) { bcost = cost_; bmv.x = omv.x + (dx); bmv.y = omv.y + (dy); } }
#define COST_MV_X4(x0, y0, x1, y1, x2, y2, x3, y3) { COST_MV_1(x0, y0) COST_MV_1(x1, y1) COST_MV_1(x2, y2) COST_MV_1(x3, y3) }
#define DIA1_ITER(mx, my) { omv.x = (mx); omv.y = (my); COST_MV_X4(0, -1, 0, 1, -1, 0, 1, 0) }
#define CROSS(start, x_max, y_max) \
        { \
            int i_ = (start); \
            if ((x_max) <= imin(mvmax.x - omv.x, omv.x - mvmin.x)) \
                for (; i_ < (x_max) - 2; i_ += 4) \
                    COST_MV_X4(i_, 0, -i_, 0, i_ + 2, 0, -i_ - 2, 0) \
            for (; i_ < (x_max); i_ += 2) \
            { \
                if (omv.x + i_ <= mvmax.x) COST_MV(omv.x + i_, omv.y) \
                if (omv.x - i_ >= mvmin.x) COST_MV(omv.x - i_, omv.y) \
            } \
            i_ = (start); \
            if ((y_max) <= imin(mvmax.y - omv.y, omv.y - mvmin.y)) \
                for (; i_ < (y_max) - 2; i_ += 4) \
                    COST_MV_X4(0, i_, 0, -i_, 0, i_ + 2, 0, -i_ - 2) \
            for (; i_ < (y_max); i_ += 2) \
            { \
                if (omv.y + i_ <= mvmax.y) COST_MV(omv.x, omv.y + i_) \
                if (omv.y - i_ >= mvmin.y) COST_MV(omv.x, omv.y - i_) \
            } \
        }
#define SAD_THRESH(v) (bcost < (((v) >> 4) * 4))      /* sizeScale[LUMA_8x8] = (8 * 8) >> 4 (motion.cpp:61,126) */
        do
        {
            ucost1 = bcost;
            DIA1_ITER(pmv.x, pmv.y)
            if (pmv.x | pmv.y)
                DIA1_ITER(0, 0)
            ucost2 = bcost;
            if ((bmv.x | bmv.y) && (bmv.x != pmv.x || bmv.y != pmv.y))
                DIA1_ITER(bmv.x, bmv.y)
            if (bcost == ucost2)
                cross_start = 3;
            omv = bmv;
            if (bcost == ucost2 && SAD_THRESH(2000))
            {
                COST_MV_X4(0, -2, -1, -1, 1, -1, -2, 0)
                COST_MV_X4(2, 0, -1, 1, 1, 1, 0, 2)
                if (bcost == ucost1 && SAD_THRESH(500))
                    break;
                if (bcost == ucost2)
                {
                    const int range = (int16_t)(merange >> 1) | 1;
                    CROSS(3, range, range)
                    COST_MV_X4(-1, -2, 1, -2, -2, -1, 2, -1)
                    COST_MV_X4(-2, 1, 2, 1, -1, 2, 1, 2)
                    if (bcost == ucost2)
                        break;
                    cross_start = range + 2;
                }
            }
            CROSS(cross_start, merange, merange >> 1)
            COST_MV_X4(-2, -2, -2, 2, 2, -2, 2, 2)
            /* hexagon grid, motion.cpp:1075-1154 */
            omv = bmv;
            int i = 1;
            do
            {
                if (4 * i > imin(imin(mvmax.x - omv.x, omv.x - mvmin.x), imin(mvmax.y - omv.y, omv.y - mvmin.y)))
                {
                    for (int j = 0; j < 16; j++)
                    {
                        or_mv mv = { omv.x + hex4[j].x * i, omv.y + hex4[j].y * i };
                        if (in_range(mv, mvmin, mvmax))
                            COST_MV(mv.x, mv.y)
                    }
                }
                else
                {
                    /* all 16 points are measured; MIN_MV tests the UNSCALED y offset against the range (motion.cpp:1100) */
                    int best = -1;
                    for (int k = 0; k < 16; k++)
                    {
                        const int mx = omv.x + hex4[k].x * i, my = omv.y + hex4[k].y * i;
                        const int cost = sad_fpel(m, mx, my) + mvcost_q(m, mx << 2, my << 2);
                        if ((omv.y + hex4[k].y >= mvmin.y) & (omv.y + hex4[k].y <= mvmax.y))
                            if (cost < bcost) { bcost = cost; best = k; }
                    }
                    if (best >= 0)
                    {
                        bmv.x = omv.x + i * hex4[best].x;
                        bmv.y = omv.y + i * hex4[best].y;
                    }
                }
            }
            while (++i <= merange >> 2);
            if (in_range(bmv, mvmin, mvmax))
                hexRefine = 1;          /* goto me_hex2 */
        }
        while (0);
#undef COST_MV
#undef COST_MV_1
#undef COST_MV_X4
#undef DIA1_ITER
#undef CROSS
#undef SAD_THRESH
    }
    else if (method == OR_STAR_SEARCH)
        star_search(m, mvmin, mvmax, merange, &bcost, &bmv);
    else if (method == OR_FULL_SEARCH)
    {   /* motion.cpp:1421-1466 with ref->isHMELowres: every vector of the range cut to [-merange, merange], raster order */
        const int r = abs(merange);
        for (int y = imax(mvmin.y, -r); y <= imin(mvmax.y, r); y++)
            for (int x = imax(mvmin.x, -r); x <= imin(mvmax.x, r); x++)
            {
                int cost = sad_fpel(m, x, y) + mvcost_q(m, x << 2, y << 2);
                if (cost < bcost) { bcost = cost; bmv.x = x; bmv.y = y; }
            }
    }
    if (hexRefine)
    {
        int c0 = COSTAT(-2, 0), c1 = COSTAT(-1, 2), c2 = COSTAT(1, 2);
        bcost <<= 3;
        if (YOK(0)) { if ((c0 << 3) + 2 < bcost) bcost = (c0 << 3) + 2; }
        if (YOK(2)) { if ((c1 << 3) + 3 < bcost) bcost = (c1 << 3) + 3; if ((c2 << 3) + 4 < bcost) bcost = (c2 << 3) + 4; }
        c0 = COSTAT(2, 0); c1 = COSTAT(1, -2); c2 = COSTAT(-1, -2);
        if (YOK(0)) { if ((c0 << 3) + 5 < bcost) bcost = (c0 << 3) + 5; }
        if (YOK(-2)) { if ((c1 << 3) + 6 < bcost) bcost = (c1 << 3) + 6; if ((c2 << 3) + 7 < bcost) bcost = (c2 << 3) + 7; }
        if (bcost & 7)
        {
            int dir = (bcost & 7) - 2;
            if (YOK(hex2[dir + 1].y))
            {
                bmv.x += hex2[dir + 1].x; bmv.y += hex2[dir + 1].y;
                for (int i = (merange >> 1) - 1;
                     i > 0 && bmv.x >= mvmin.x && bmv.x <= mvmax.x && bmv.y >= mvmin.y && bmv.y <= mvmax.y; i--)
                {
                    c0 = COSTAT(hex2[dir + 0].x, hex2[dir + 0].y);
                    c1 = COSTAT(hex2[dir + 1].x, hex2[dir + 1].y);
                    c2 = COSTAT(hex2[dir + 2].x, hex2[dir + 2].y);
                    bcost &= ~7;
                    if (YOK(hex2[dir + 0].y)) { if ((c0 << 3) + 1 < bcost) bcost = (c0 << 3) + 1; }
                    if (YOK(hex2[dir + 1].y)) { if ((c1 << 3) + 2 < bcost) bcost = (c1 << 3) + 2; }
                    if (YOK(hex2[dir + 2].y)) { if ((c2 << 3) + 3 < bcost) bcost = (c2 << 3) + 3; }
                    if (!(bcost & 7))
                        break;
                    dir += (bcost & 7) - 2;
                    dir = mod6m1[dir + 1];
                    bmv.x += hex2[dir + 1].x; bmv.y += hex2[dir + 1].y;
                }
            }
        }
        bcost >>= 3;
        /* square refine, motion.cpp:950-967 */
        int dir = 0, c3;
        c0 = COSTAT(0, -1); c1 = COSTAT(0, 1); c2 = COSTAT(-1, 0); c3 = COSTAT(1, 0);
        if (YOK(-1)) { if (c0 < bcost) { bcost = c0; dir = 1; } }
        if (YOK(1))  { if (c1 < bcost) { bcost = c1; dir = 2; } }
        if (c2 < bcost) { bcost = c2; dir = 3; }
        if (c3 < bcost) { bcost = c3; dir = 4; }
        c0 = COSTAT(-1, -1); c1 = COSTAT(-1, 1); c2 = COSTAT(1, -1); c3 = COSTAT(1, 1);
        if (YOK(-1)) { if (c0 < bcost) { bcost = c0; dir = 5; } }
        if (YOK(1))  { if (c1 < bcost) { bcost = c1; dir = 6; } }
        if (YOK(-1)) { if (c2 < bcost) { bcost = c2; dir = 7; } }
        if (YOK(1))  { if (c3 < bcost) { bcost = c3; dir = 8; } }
        bmv.x += square1[dir].x; bmv.y += square1[dir].y;
    }
#undef YOK
#undef COSTAT
    if (bprecost < bcost) { bmv = bestpre; bcost = bprecost; }
    else { bmv.x <<= 2; bmv.y <<= 2; }

    if (!bcost)
        bcost = mvcost_q(m, bmv.x, bmv.y);
    else
    {
        int bdir = 0;
        for (int i = 1; i <= 4; i++)
        {
            int qx = bmv.x + square1[i].x * 2, qy = bmv.y + square1[i].y * 2;
            if ((qy < qmin.y) | (qy > qmax.y)) continue;
            int cost = qpel_cost(m, qx, qy, 0) + mvcost_q(m, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += square1[bdir].x * 2; bmv.y += square1[bdir].y * 2;
        bcost = qpel_cost(m, bmv.x, bmv.y, 1) + mvcost_q(m, bmv.x, bmv.y);
        bdir = 0;
        for (int i = 1; i <= 4; i++)
        {
            int qx = bmv.x + square1[i].x, qy = bmv.y + square1[i].y;
            if ((qy < qmin.y) | (qy > qmax.y)) continue;
            int cost = qpel_cost(m, qx, qy, 1) + mvcost_q(m, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmv.x += square1[bdir].x; bmv.y += square1[bdir].y;
    }
    *out = bmv;
    return bcost;
}

static int motion_estimate(or_me* m, or_mv mvmin, or_mv mvmax, or_mv qmvp, or_mv* out)
{
    return motion_estimate_ex(m, mvmin, mvmax, qmvp, 16, OR_HEX_SEARCH, out);
}

/* is cuY the first row a search visits in its slice?  Whole frame: the bottom row (slicetype.cpp:4050-4059);
 * cooperative slices: the bottom row of each slice, the last slice running to the frame's end (:3957-3968) */
static int or_last_row(const or_geom* g, int cuY)
{
    if (g->rowsPerSlice <= 0) return cuY == g->bh - 1;
    const int ns = g->bh / g->rowsPerSlice;
    int si = cuY / g->rowsPerSlice;
    if (si > ns - 1) si = ns - 1;
    const int lastY = si == ns - 1 ? g->bh - 1 : (si + 1) * g->rowsPerSlice - 1;
    return cuY == lastY;
}

/* search half of estimateCUCost over a whole frame (slicetype.cpp:4050-4059, 4103-4183) */
void or_search_list(const or_geom* g, const or_pixel* fencPlane0, const or_pixel* const refPlanes[4],
                    const uint16_t* mvcost, int bBidir, int32_t* mvs, int32_t* mvCosts, int32_t* skipCount)
{
    int skips = 0;
    or_me m;
    m.g = g; m.stride = g->stride; m.mvcost = mvcost;
    const int bw = g->bw, bh = g->bh;
    for (int cuY = bh - 1; cuY >= 0; cuY--)
    {
        const int lastRow = or_last_row(g, cuY);
        for (int cuX = bw - 1; cuX >= 0; cuX--)
        {
            const int cuXY = cuX + cuY * bw;
            const int64_t pel = 8 * cuX + (int64_t)8 * cuY * g->stride;
            for (int y = 0; y < 8; y++) memcpy(m.fenc + 8 * y, fencPlane0 + pel + (int64_t)y * g->stride, 8 * sizeof(or_pixel));
            for (int i = 0; i < 4; i++) m.ref[i] = refPlanes[i] + pel;
            or_mv mvmin = { -cuX * 8 - 8, -cuY * 8 - 8 };
            or_mv mvmax = { (bw - cuX - 1) * 8 + 8, (bh - cuY - 1) * 8 + 8 };

            or_mv mvc[4]; int numc = 0;
#define MVC(idx) { mvc[numc].x = mvs[2 * (idx)]; mvc[numc].y = mvs[2 * (idx) + 1]; numc++; }
            if (cuX < bw - 1) MVC(cuXY + 1);
            if (!lastRow)
            {
                MVC(cuXY + bw);
                if (cuX > 0) MVC(cuXY + bw - 1);
                if (cuX < bw - 1) MVC(cuXY + bw + 1);
            }
#undef MVC
            or_mv mvp = { 0, 0 };
            int skipCost = 0x7fffffff;
            if (numc)
            {
                int mvpcost = OR_COST_MAX;
                or_pixel buf[64];
                for (int i = 0; i < numc; i++)
                {
                    int st;
                    const or_pixel* src = lowres_mc(&m, mvc[i].x, mvc[i].y, buf, &st);
                    int cost = or_satd8x8(m.fenc, 8, src, st);
                    if (cost < mvpcost) { mvpcost = cost; mvp = mvc[i]; }
                    if (!(mvp.x | mvp.y) && bBidir)
                        skipCost = cost;
                }
            }
            or_mv best;
            int fencCost = motion_estimate(&m, mvmin, mvmax, mvp, &best);
            if (skipCost < 64 && skipCost < fencCost && bBidir)
            {
                fencCost = skipCost;
                best.x = best.y = 0;
                skips++;
            }
            mvs[2 * cuXY] = best.x; mvs[2 * cuXY + 1] = best.y;
            mvCosts[cuXY] = fencCost;
        }
    }
    if (skipCount) *skipCount = skips;
}

/* --hme: the search half of estimateCUCost for one level (slicetype.cpp:3942-3955, 4040-4048, 4083-4183) */
void or_search_list_hme(const or_geom* g, int level, const or_pixel* fencPlane0, const or_pixel* const refPlanes[4],
                        const uint16_t* mvcost, int bBidir, int method, int merange,
                        const int32_t* hmeMvs, const int32_t* hmeMvCosts,
                        int32_t* mvs, int32_t* mvCosts, int32_t* skipCount)
{
    int skips = 0;
    or_me m;
    const int hme = level == 0;
    const int bw = hme ? g->bw4 : g->bw, bh = hme ? g->bh4 : g->bh;
    const int stride = hme ? g->stride4 : g->stride;
    m.g = g; m.stride = stride; m.mvcost = mvcost;
    /* No cooperative slices here: with them the reference's level-1 search of one slice reads level-0 vectors that another
     * slice's worker may not have written yet (the two levels are cut at different rows, :3942-3968) -- uninitialised
     * vectors, observed to crash the reference.  The GPU path refuses --hme together with active lookahead slices. */
    for (int cuY = bh - 1; cuY >= 0; cuY--)
    {
        const int lastRow = cuY == bh - 1;
        for (int cuX = bw - 1; cuX >= 0; cuX--)
        {
            const int cuXY = cuX + cuY * bw;
            const int cuXY_4x4 = (cuX / 2) + (cuY / 2) * bw / 2;       /* :4088, as written */
            const int64_t pel = 8 * cuX + (int64_t)8 * cuY * stride;
            for (int y = 0; y < 8; y++) memcpy(m.fenc + 8 * y, fencPlane0 + pel + (int64_t)y * stride, 8 * sizeof(or_pixel));
            for (int i = 0; i < 4; i++) m.ref[i] = refPlanes[i] + pel;
            or_mv mvmin = { -cuX * 8 - 8, -cuY * 8 - 8 };
            or_mv mvmax = { (bw - cuX - 1) * 8 + 8, (bh - cuY - 1) * 8 + 8 };

            or_mv mvc[5]; int numc = 0;
#define MVC(idx) { mvc[numc].x = mvs[2 * (idx)]; mvc[numc].y = mvs[2 * (idx) + 1]; numc++; }
            if (cuX < bw - 1) MVC(cuXY + 1);
            if (!lastRow)
            {
                MVC(cuXY + bw);
                if (cuX > 0) MVC(cuXY + bw - 1);
                if (cuX < bw - 1) MVC(cuXY + bw + 1);
            }
#undef MVC
            if (!hme && hmeMvs && hmeMvCosts[cuXY_4x4] > 0)
            {
                mvc[numc].x = hmeMvs[2 * cuXY_4x4] * 2; mvc[numc].y = hmeMvs[2 * cuXY_4x4 + 1] * 2;
                numc++;
            }
            or_mv mvp = { 0, 0 };
            int skipCost = 0x7fffffff;
            if (numc)
            {
                int mvpcost = OR_COST_MAX;
                or_pixel buf[64];
                for (int i = 0; i < numc; i++)
                {
                    int st;
                    const or_pixel* src = lowres_mc(&m, mvc[i].x, mvc[i].y, buf, &st);
                    int cost = or_satd8x8(m.fenc, 8, src, st);
                    if (cost < mvpcost) { mvpcost = cost; mvp = mvc[i]; }
                    if (!(mvp.x | mvp.y) && bBidir)
                        skipCost = cost;
                }
            }
            or_mv best;
            int fencCost = motion_estimate_ex(&m, mvmin, mvmax, mvp, merange, method, &best);
            if (skipCost < 64 && skipCost < fencCost && bBidir)
            {
                fencCost = skipCost;
                best.x = best.y = 0;
                skips++;
            }
            mvs[2 * cuXY] = best.x; mvs[2 * cuXY + 1] = best.y;
            mvCosts[cuXY] = fencCost;
        }
    }
    if (skipCount) *skipCount = skips;
}

/* cost half of estimateCUCost + frame sums (slicetype.cpp:4187-4248) */
void or_frame_cost(const or_geom* g, const or_pixel* fencPlane0,
                   const or_pixel* const ref0Planes[4], const or_pixel* const ref1Planes[4],
                   const int32_t* mvs0, const int32_t* mvCosts0,
                   const int32_t* mvs1, const int32_t* mvCosts1,
                   const int32_t* intraCost, const int32_t* invQ,
                   uint16_t* lowresCosts, int32_t* rowSatds,
                   int64_t* costEstOut, int64_t* costEstAqOut, int32_t* intraMbsOut)
{
    const int bw = g->bw, bh = g->bh, bBidir = ref1Planes != 0;
    int64_t costEst = 0, costEstAq = 0; int intraMbs = 0;
    or_me m0, m1;
    m0.stride = m1.stride = g->stride;
    for (int cuY = bh - 1; cuY >= 0; cuY--)
    {
        rowSatds[cuY] = 0;
        for (int cuX = bw - 1; cuX >= 0; cuX--)
        {
            const int cuXY = cuX + cuY * bw;
            const int64_t pel = 8 * cuX + (int64_t)8 * cuY * g->stride;
            int bcost = OR_COST_MAX, listused = 0;
            if (mvCosts0[cuXY] < bcost) { bcost = mvCosts0[cuXY]; listused = 1; }
            if (bBidir)
            {
                if (mvCosts1[cuXY] < bcost) { bcost = mvCosts1[cuXY]; listused = 2; }
                or_pixel fenc[64], b0[64], b1[64], avg[64];
                for (int y = 0; y < 8; y++) memcpy(fenc + 8 * y, fencPlane0 + pel + (int64_t)y * g->stride, 8 * sizeof(or_pixel));
                for (int i = 0; i < 4; i++) { m0.ref[i] = ref0Planes[i] + pel; m1.ref[i] = ref1Planes[i] + pel; }
                int s0, s1;
                const or_pixel* p0 = lowres_mc(&m0, mvs0[2 * cuXY], mvs0[2 * cuXY + 1], b0, &s0);
                const or_pixel* p1 = lowres_mc(&m1, mvs1[2 * cuXY], mvs1[2 * cuXY + 1], b1, &s1);
                for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) avg[y * 8 + x] = (or_pixel)((p0[y * s0 + x] + p1[y * s1 + x] + 1) >> 1);
                int bicost = or_satd8x8(fenc, 8, avg, 8);
                if (bicost < bcost) { bcost = bicost; listused = 3; }
                p0 = m0.ref[0]; p1 = m1.ref[0];
                for (int y = 0; y < 8; y++) for (int x = 0; x < 8; x++) avg[y * 8 + x] = (or_pixel)((p0[(int64_t)y * g->stride + x] + p1[(int64_t)y * g->stride + x] + 1) >> 1);
                bicost = or_satd8x8(fenc, 8, avg, 8);
                if (bicost < bcost) { bcost = bicost; listused = 3; }
                bcost += 4;
            }
            else
            {
                bcost += 4;
                if (intraCost[cuXY] < bcost) { bcost = intraCost[cuXY]; listused = 0; }
            }
            int scored = (cuX > 0 && cuX < bw - 1 && cuY > 0 && cuY < bh - 1) || bw <= 2 || bh <= 2;
            int bcostAq = (scored && invQ) ? ((bcost * invQ[cuXY] + 128) >> 8) : bcost;
            if (scored)
            {
                costEst += bcost; costEstAq += bcostAq;
                if (!listused && !bBidir) intraMbs++;
            }
            rowSatds[cuY] += bcostAq;
            lowresCosts[cuXY] = (uint16_t)(imin(bcost, OR_LOWRES_COST_MASK) | (listused << OR_LOWRES_COST_SHIFT));
        }
    }
    *costEstOut = costEst; *costEstAqOut = costEstAq; *intraMbsOut = intraMbs;
}

/* ------------------------------------------------------------------ weighted prediction */

/* weight_pp_c over whole padded planes (pixel.cpp:518-541, slicetype.cpp:833-842,966-976) */
void or_weight_planes(const or_geom* g, const or_pixel* refBuf, or_pixel* dstBuf, int nPlanes,
                      int scale, int denom, int offsetIn)
{
    const int correction = 14 - OR_DEPTH;
    const int offset = offsetIn << (OR_DEPTH - 8);
    const int round = (denom ? 1 << (denom - 1) : 0) << correction;
    const int shift = denom + correction;
    const int64_t n = g->planeSize * nPlanes;
    for (int64_t i = 0; i < n; i++)
    {
        int16_t val = (int16_t)(refBuf[i] << correction);
        int v = ((scale * val + round) >> shift) + offset;
        dstBuf[i] = (or_pixel)(v < 0 ? 0 : (v > OR_MAXV ? OR_MAXV : v));
    }
}

/* slicetype.cpp:845-858 */
uint32_t or_weight_cost_luma(const or_geom* g, const or_pixel* fencPlane0, const or_pixel* refPlane0,
                             const int32_t* intraCost)
{
    uint32_t cost = 0; int mb = 0;
    for (int y = 0; y < g->h; y += 8)
        for (int x = 0; x < g->w; x += 8, mb++)
        {
            int64_t off = (int64_t)y * g->stride + x;
            int satd = or_satd8x8(refPlane0 + off, g->stride, fencPlane0 + off, g->stride);
            cost += (uint32_t)imin(satd, intraCost[mb]);
        }
    return cost;
}

/* slicetype.cpp:879-980 */
int or_weights_analyse(const or_geom* g, const or_pixel* fencPlane0, const or_pixel* refBuf,
                       const int32_t* intraCost, const uint64_t fenc_ssd, const uint64_t fenc_sum,
                       const uint64_t ref_ssd, const uint64_t ref_sum,
                       or_pixel* wbuf, int* scaleOut, int* denomOut, int* offsetOut, double* costDelta)
{
    const float epsilon = 1.f / 128.f;
    float guessScale, fencMean, refMean;
    if (fenc_ssd && ref_ssd) guessScale = sqrtf((float)fenc_ssd / ref_ssd);
    else guessScale = 1.0f;
    fencMean = (float)fenc_sum / (g->h * g->w) / (1 << (OR_DEPTH - 8));
    refMean = (float)ref_sum / (g->h * g->w) / (1 << (OR_DEPTH - 8));
    if (fabsf(refMean - fencMean) < 0.5f && fabsf(1.f - guessScale) < epsilon)
        return 0;

    int minoff = 0, minscale, mindenom, found = 0;
    unsigned int minscore, origscore;
    {   /* WeightParam::setFromWeightAndOffset(w, 0, 7, true) (common/slice.h:304-316) */
        int w = (int)(guessScale * 128 + 0.5f), d = 7;
        while (d > 0 && w > 127) { d--; w >>= 1; }
        w = imin(w, 127);
        mindenom = d; minscale = w;
    }
    const or_pixel* refPlane0 = refBuf + g->padOffset;
    origscore = minscore = or_weight_cost_luma(g, fencPlane0, refPlane0, intraCost);
    if (!minscore)
        return 0;
    int curScale = minscale;
    int curOffset = (int)(fencMean - refMean * curScale / (1 << mindenom) + 0.5f);
    if (curOffset < -128 || curOffset > 127)
    {
        curOffset = imax(-128, imin(127, curOffset));
        curScale = (int)((1 << mindenom) * (fencMean - curOffset) / refMean + 0.5f);
        curScale = imax(0, imin(127, curScale));
    }
    or_weight_planes(g, refBuf, wbuf, 1, curScale, mindenom, curOffset);
    unsigned int s = or_weight_cost_luma(g, fencPlane0, wbuf + g->padOffset, intraCost);
    if (s < minscore) { minscore = s; minscale = curScale; minoff = curOffset; found = 1; }
    if (mindenom > 0 && !(minscale & 1))
    {
        int idx = 0;
        while (!((minscale >> idx) & 1)) idx++;      /* CTZ */
        int shift = imin(idx, mindenom);
        mindenom -= shift;
        minscale >>= shift;
    }
    if (!found || (minscale == (1 << mindenom) && minoff == 0) || (float)minscore / origscore > 0.998f)
        return 0;
    *costDelta = (double)(minscore / origscore);     /* integer division, as the reference (slicetype.cpp:964) */
    or_weight_planes(g, refBuf, wbuf, 4, minscale, mindenom, minoff);
    *scaleOut = minscale; *denomOut = mindenom; *offsetOut = minoff;
    return 1;
}

/* ------------------------------------------------------------------ cuTree */

static inline void clip_add(uint16_t* s, int x)
{
    int v = *s + x;
    *s = (uint16_t)(v < 65535 ? v : 65535);
}

/* slicetype.cpp:3502-3604 + pixel.cpp:931-957 */
void or_cutree_propagate(const or_geom* g, const int32_t* intraCost, const uint16_t* lowresCosts,
                         const int32_t* invQ, const int32_t* mvs0, const int32_t* mvs1,
                         uint16_t* propagateB, uint16_t* refCost0, uint16_t* refCost1,
                         int referenced, int bipredWeight, double fpsFactor)
{
    const int bw = g->bw, bh = g->bh;
    uint16_t* refCosts[2] = { refCost0, refCost1 };
    const int32_t* mvsL[2] = { mvs0, mvs1 };
    const int bipredWeights[2] = { bipredWeight, 64 - bipredWeight };
    if (!referenced)
        memset(propagateB, 0, bw * sizeof(uint16_t));
    const uint16_t* propIn = propagateB;
    const double fps = fpsFactor / 256;
    for (int by = 0; by < bh; by++)
    {
        for (int bx = 0; bx < bw; bx++)
        {
            const int cu = by * bw + bx;
            int intra = intraCost[cu];
            int inter = imin(intraCost[cu], lowresCosts[cu] & OR_LOWRES_COST_MASK);
            double propagateIntra = intra * (invQ ? invQ[cu] : 256);
            double propagateAmount = (double)propIn[bx] + propagateIntra * fps;
            double propagateNum = (double)(intra - inter);
            double propagateDenom = (double)intra;
            int amount = (int)(propagateAmount * propagateNum / propagateDenom + 0.5);
            if (amount <= 0) continue;
            int lists_used = lowresCosts[cu] >> OR_LOWRES_COST_SHIFT;
            for (int list = 0; list < 2; list++)
            {
                if (!((lists_used >> list) & 1)) continue;
                int listamount = amount;
                if (lists_used == 3)
                    listamount = (listamount * bipredWeights[list] + 32) >> 6;
                int x = mvsL[list][2 * cu], y = mvsL[list][2 * cu + 1];
                if (!x && !y) { clip_add(&refCosts[list][cu], listamount); continue; }
                int cux = (x >> 5) + bx, cuy = (y >> 5) + by;
                int idx0 = cux + cuy * bw, idx1 = idx0 + 1, idx2 = idx0 + bw, idx3 = idx0 + bw + 1;
                x &= 31; y &= 31;
                int w0 = (32 - y) * (32 - x), w1 = (32 - y) * x, w2 = y * (32 - x), w3 = y * x;
                if (cux < bw - 1 && cuy < bh - 1 && cux >= 0 && cuy >= 0)
                {
                    clip_add(&refCosts[list][idx0], (listamount * w0 + 512) >> 10);
                    clip_add(&refCosts[list][idx1], (listamount * w1 + 512) >> 10);
                    clip_add(&refCosts[list][idx2], (listamount * w2 + 512) >> 10);
                    clip_add(&refCosts[list][idx3], (listamount * w3 + 512) >> 10);
                }
                else
                {
                    if (cux < bw && cuy < bh && cux >= 0 && cuy >= 0)
                        clip_add(&refCosts[list][idx0], (listamount * w0 + 512) >> 10);
                    if (cux + 1 < bw && cuy < bh && cux + 1 >= 0 && cuy >= 0)
                        clip_add(&refCosts[list][idx1], (listamount * w1 + 512) >> 10);
                    if (cux < bw && cuy + 1 < bh && cux >= 0 && cuy + 1 >= 0)
                        clip_add(&refCosts[list][idx2], (listamount * w2 + 512) >> 10);
                    if (cux + 1 < bw && cuy + 1 < bh && cux + 1 >= 0 && cuy + 1 >= 0)
                        clip_add(&refCosts[list][idx3], (listamount * w3 + 512) >> 10);
                }
            }
        }
        if (referenced)
            propIn += bw;
    }
}

/* slicetype.cpp:3784-3796 */
void or_cutree_finish(const or_geom* g, const int32_t* intraCost, const int32_t* invQ,
                      const uint16_t* propagateCost, const double* qpAqOffset, double* qpCuTreeOffset,
                      int fpsFactorFix8, double weightdelta, double cuTreeStrength)
{
    if (g->qg8)
    {
        /* slicetype.cpp:3764-3782; invQ = invQscaleFactor8x8 */
        const int fs = 2 * g->bw;
        for (int cuY = 0; cuY < g->bh; cuY++)
            for (int cuX = 0; cuX < g->bw; cuX++)
            {
                const int cu = cuX + cuY * g->bw;
                int intracost = ((intraCost[cu]) / 4 * invQ[cu] + 128) >> 8;
                if (intracost)
                {
                    int propagate = ((propagateCost[cu]) / 4 * fpsFactorFix8 + 128) >> 8;
                    double log2_ratio = log2((double)(intracost + propagate)) - log2((double)intracost) + weightdelta;
                    const int i = cuX * 2 + cuY * g->bw * 4;
                    qpCuTreeOffset[i] = qpAqOffset[i] - cuTreeStrength * (log2_ratio);
                    qpCuTreeOffset[i + 1] = qpAqOffset[i + 1] - cuTreeStrength * (log2_ratio);
                    qpCuTreeOffset[i + fs] = qpAqOffset[i + fs] - cuTreeStrength * (log2_ratio);
                    qpCuTreeOffset[i + fs + 1] = qpAqOffset[i + fs + 1] - cuTreeStrength * (log2_ratio);
                }
            }
        return;
    }
    for (int i = 0; i < g->ncu; i++)
    {
        int intracost = (intraCost[i] * invQ[i] + 128) >> 8;
        if (intracost)
        {
            int propagate = (propagateCost[i] * fpsFactorFix8 + 128) >> 8;
            double log2_ratio = log2((double)(intracost + propagate)) - log2((double)intracost) + weightdelta;
            qpCuTreeOffset[i] = qpAqOffset[i] - cuTreeStrength * log2_ratio;
        }
    }
}

/* qg-size 8: the lowres block's factor is the mean of its four 8x8 factors (slicetype.cpp:656-670) */
void or_invq8x8(const or_geom* g, const int32_t* invQ, int32_t* invQ8)
{
    const int fs = 2 * g->bw;
    for (int cuY = 0; cuY < g->bh; cuY++)
        for (int cuX = 0; cuX < g->bw; cuX++)
        {
            const int i = cuX * 2 + cuY * g->bw * 4;
            invQ8[cuX + cuY * g->bw] = (invQ[i] + invQ[i + 1] + invQ[i + fs] + invQ[i + fs + 1]) / 4;
        }
}

static double block_qp_offset(const or_geom* g, const double* qp, int cux, int cuy)
{
    if (!g->qg8) return qp[cux + cuy * g->bw];
    const int fs = 2 * g->bw, i = cux * 2 + cuy * g->bw * 4;
    return (qp[i] + qp[i + 1] + qp[i + fs] + qp[i + fs + 1]) / 4;
}

/* slicetype.cpp:3847-3878 */
int64_t or_frame_cost_recalc(const or_geom* g, const uint16_t* lowresCosts, const double* qpOffset, int32_t* rowSatds)
{
    int64_t score = 0;
    for (int cuy = g->bh - 1; cuy >= 0; cuy--)
    {
        rowSatds[cuy] = 0;
        for (int cux = g->bw - 1; cux >= 0; cux--)
        {
            int cu = cux + cuy * g->bw;
            int cuCost = lowresCosts[cu] & OR_LOWRES_COST_MASK;
            cuCost = (cuCost * or_exp2fix8(block_qp_offset(g, qpOffset, cux, cuy)) + 128) >> 8;
            rowSatds[cuy] += cuCost;
            if ((cuy > 0 && cuy < g->bh - 1 && cux > 0 && cux < g->bw - 1) || g->bw <= 2 || g->bh <= 2)
                score += cuCost;
        }
    }
    return score;
}

/* the VBV half of getEstimatedPictureCost (slicetype.cpp:1387-1436); qpOffset may be NULL */
void or_vbv_rows(const or_geom* g, const uint16_t* lowresCosts, const int32_t* intraCost, const double* qpOffset,
                 int scale, int pirStart, int pirEnd, int nRows, uint32_t* satdForVbv, uint32_t* intraSatdForVbv,
                 uint16_t* lowresCostForRc, int32_t* intraCostScaled)
{
    for (int i = 0; i < nRows; i++) satdForVbv[i] = intraSatdForVbv[i] = 0;
    for (int row = 0; row < nRows; row++)
    {
        int lowresRow = row * scale;
        for (int cnt = 0; cnt < scale && lowresRow < g->bh; lowresRow++, cnt++)
        {
            uint32_t sum = 0, intraSum = 0;
            int diff = 0;
            int idx = lowresRow * g->bw;
            for (int col = 0; col < g->bw; col++, idx++)
            {
                uint16_t c = lowresCosts[idx] & OR_LOWRES_COST_MASK;
                int32_t ic = intraCost[idx];
                if (qpOffset)
                {
                    double q = block_qp_offset(g, qpOffset, col, lowresRow);
                    c = (uint16_t)((c * or_exp2fix8(q) + 128) >> 8);
                    ic = (ic * or_exp2fix8(q) + 128) >> 8;
                }
                if (pirStart >= 0)
                    for (int x = pirStart; x <= pirEnd; x++)
                        diff += ic - c;
                lowresCostForRc[idx] = c;
                intraCostScaled[idx] = ic;
                sum += c;
                intraSum += ic;
            }
            satdForVbv[row] += sum;
            satdForVbv[row] += diff;
            intraSatdForVbv[row] += intraSum;
        }
    }
}
