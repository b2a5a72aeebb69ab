/*
 * orb_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see orb_oracle.h): OpenCV primitive
 * restatements + the ORBextractor pipeline, written as straightforward sequential C.
 *
 * Reference citations use S/ = /root/reference/oRB_SLAM2_Android/src/main/jni/ORB_SLAM2/src/.
 * The OpenCV primitives are not in /root/reference (OpenCV 2.4.9 Android SDK, un-vendored,
 * J/Android.mk:30); they are restated from the published algorithms and pinned bit-for-bit
 * against cv2 4.13 by tests/test_oracle_primitives.py.
 */
#include "orb_oracle.h"
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================================= */
/* scalar helpers                                                                          */
/* ======================================================================================= */

/* cvRound(float): SSE cvtss2si under the default rounding mode = round half to even. */
int orc_round(float v) { return (int)lrintf(v); }

/* cv::fastAtan2 (OpenCV core/mathfuncs): 7th-order odd polynomial in min/max, octant fix-ups.
 * Call site S/ORBextractor.cc:108.  Every operation is a separately rounded fp32 op. */
float orc_fast_atan2(float y, float x)
{
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    volatile float ax = fabsf(x), ay = fabsf(y);
    volatile float a, c, c2, t;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
        a = t * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
        t = t * c;
        a = 90.f - t;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* glibc >= 2.28 sinf/cosf/sincosf (the ARM optimized-routines algorithm): reduce by pi/2 in
 * double, evaluate a double polynomial, round once to float.  The reference calls
 * cos(float)/sin(float) at S/ORBextractor.cc:118; on x86-64 that resolves to this libm code.
 * tests/test_oracle_primitives.py checks this restatement against libm for EVERY float in
 * [0, 2*pi] (the only domain the extractor can produce). */
static float sc_poly(double x, double x2, int n, int negcos)
{
    static const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5,
                        C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    static const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7,
                        S3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double s1 = S2 + x2 * S3;
        double x7 = x3 * x2;
        double s = x + x3 * S1;
        return (float)(s + x7 * s1);
    } else {
        double sg = negcos ? -1.0 : 1.0;
        double x4 = x2 * x2;
        double c2 = sg * C3 + x2 * (sg * C4);
        double c1 = sg * C0 + x2 * (sg * C1);
        double x6 = x4 * x2;
        double c = c1 + x4 * (sg * C2);
        return (float)(c + x6 * c2);
    }
}

void orc_sincosf(float y, float *sp, float *cp)
{
    static const double HPI_INV = 0x1.45F306DC9C883p+23, HPI = 0x1.921FB54442D18p0;
    static const double SIGN[4] = {1.0, -1.0, -1.0, 1.0};
    double x = y;
    uint32_t u;
    memcpy(&u, &y, 4);
    uint32_t top = (u >> 20) & 0x7ff;
    if (top < 0x3f4) {            /* |y| < pi/4 */
        if (top < 0x398) {        /* |y| < 2^-12 */
            *sp = y;
            *cp = 1.0f;
            return;
        }
        double x2 = x * x;
        *sp = sc_poly(x, x2, 0, 0);
        *cp = sc_poly(x, x2, 1, 0);
        return;
    }
    /* valid for |y| < 120, far beyond the [0, 2*pi] the extractor produces */
    double r = x * HPI_INV;
    int n = ((int32_t)r + 0x800000) >> 24;
    x = x - n * HPI;
    double s = SIGN[n & 3];
    int neg = (n & 2) != 0;
    *sp = sc_poly(x * s, x * x, n, neg);
    *cp = sc_poly(x * s, x * x, n ^ 1, neg);
}

/* ======================================================================================= */
/* cv::resize(..., INTER_LINEAR) for 8UC1 (call site S/ORBextractor.cc:1157)               */
/* ======================================================================================= */
/* Fixed-point bilinear: 11-bit coefficients, horizontal pass into int, vertical pass with
 * the (>>4, *beta >>16, +2 >>2) rounding of OpenCV's 8-bit linear vertical resizer. */
static short coef11(float c) /* saturate_cast<short>(c * 2048) */
{
    int v = orc_round(c * 2048.f);
    if (v > 32767) v = 32767;
    if (v < -32768) v = -32768;
    return (short)v;
}

void orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride,
                          uint8_t *dst, int dw, int dh, int dstride)
{
    double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
    double scale_x = 1. / inv_scale_x, scale_y = 1. / inv_scale_y;
    int *xofs = (int *)malloc(sizeof(int) * dw);
    short *alpha = (short *)malloc(sizeof(short) * 2 * dw);
    int *row0 = (int *)malloc(sizeof(int) * dw), *row1 = (int *)malloc(sizeof(int) * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = coef11(1.f - fx);
        alpha[2 * dx + 1] = coef11(fx);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = coef11(1.f - fy), b1 = coef11(fy);
        int y0 = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);
        int y1 = sy + 1 < 0 ? 0 : (sy + 1 >= sh ? sh - 1 : sy + 1);
        const uint8_t *s0 = src + (size_t)y0 * sstride, *s1 = src + (size_t)y1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx];
            int sx1 = sx + 1 < sw ? sx + 1 : sx; /* weight is 0 there */
            row0[dx] = s0[sx] * alpha[2 * dx] + s0[sx1] * alpha[2 * dx + 1];
            row1[dx] = s1[sx] * alpha[2 * dx] + s1[sx1] * alpha[2 * dx + 1];
        }
        uint8_t *d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            d[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs); free(alpha); free(row0); free(row1);
}

/* cv::copyMakeBorder(..., BORDER_REFLECT_101) (S/ORBextractor.cc:1159-1164): dst has
 * (w+2b) x (h+2b) pixels; index -i maps to +i, index (n-1)+i maps to (n-1)-i. */
static int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * (n - 1) - p;
    }
    return p;
}

void orc_copy_make_border_reflect101(const uint8_t *src, int w, int h, int sstride,
                                     uint8_t *dst, int dstride, int border)
{
    for (int y = 0; y < h + 2 * border; y++) {
        const uint8_t *s = src + (size_t)reflect101(y - border, h) * sstride;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w + 2 * border; x++) d[x] = s[reflect101(x - border, w)];
    }
}

/* ======================================================================================= */
/* cv::FAST (TYPE_9_16) with optional 3x3 non-max suppression (S/ORBextractor.cc:827,831)   */
/* ======================================================================================= */
/* Bresenham circle of radius 3, clockwise from 12 o'clock (dx,dy). */
static const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* Corner score = the largest t for which the pixel is still a FAST-9 corner:
 * max over the 16 arcs of 9 contiguous ring pixels of min |v - ring| (one polarity at a
 * time), minus 1.  Returns 0 if the pixel is not a corner at `threshold`. */
static int fast_score(const uint8_t *p, const int *off, int threshold)
{
    const int v = p[0];
    int d[16], lo2[16], hi2[16], lo4[16], hi4[16];
    for (int k = 0; k < 16; k++) d[k] = (int)p[off[k]] - v;
    /* sliding minimum / maximum over the windows of 9 of the circular ring, by doubling: 2, 4, 8, +1 */
    for (int k = 0; k < 16; k++) {
        const int e = d[(k + 1) & 15];
        lo2[k] = d[k] < e ? d[k] : e;
        hi2[k] = d[k] > e ? d[k] : e;
    }
    for (int k = 0; k < 16; k++) {
        const int a = lo2[(k + 2) & 15], b = hi2[(k + 2) & 15];
        lo4[k] = lo2[k] < a ? lo2[k] : a;
        hi4[k] = hi2[k] > b ? hi2[k] : b;
    }
    int bright = -256, dark = 256;      /* max over arcs of min(d);  min over arcs of max(d) */
    for (int k = 0; k < 16; k++) {
        int lo = lo4[k] < lo4[(k + 4) & 15] ? lo4[k] : lo4[(k + 4) & 15];
        int hi = hi4[k] > hi4[(k + 4) & 15] ? hi4[k] : hi4[(k + 4) & 15];
        const int e = d[(k + 8) & 15];
        if (e < lo) lo = e;
        if (e > hi) hi = e;
        if (lo > bright) bright = lo;
        if (hi < dark) dark = hi;
    }
    const int m = bright > -dark ? bright : -dark;
    return m > threshold ? m - 1 : 0;
}

int orc_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold, int nms,
                 orc_keypoint *out, int cap)
{
    int n = 0;
    if (w < 7 || h < 7) return 0;
    if (threshold < 0) threshold = 0;
    if (threshold > 255) threshold = 255;
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = RING_DY[k] * stride + RING_DX[k];
    /* classification table: 1 = darker than centre - t, 2 = brighter than centre + t */
    uint8_t tab[512];
    for (int i = -255; i <= 255; i++) tab[i + 255] = (uint8_t)(i < -threshold ? 1 : (i > threshold ? 2 : 0));
    /* score map over the whole window; 0 = not a corner (also outside the 3-px margin).
     * Cell-sized windows (the extractor's case) use a stack buffer. */
    uint8_t stackbuf[72 * 72 * 2];
    const size_t npx = (size_t)w * h;
    uint8_t *score = npx * 2 <= sizeof(stackbuf) ? stackbuf : (uint8_t *)malloc(npx * 2);
    uint8_t *is_corner = score + npx;
    memset(score, 0, npx * 2);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; x++) {
            const uint8_t *p = row + x;
            const uint8_t *tb = tab + 255 - p[0];
            /* a 9-arc of one polarity contains a pixel of every opposite pair: AND the pair classes */
            int d = tb[p[off[0]]] | tb[p[off[8]]];
            if (!d) continue;
            d &= tb[p[off[4]]] | tb[p[off[12]]];
            d &= tb[p[off[2]]] | tb[p[off[10]]];
            d &= tb[p[off[6]]] | tb[p[off[14]]];
            if (!d) continue;
            d &= tb[p[off[1]]] | tb[p[off[9]]];
            d &= tb[p[off[3]]] | tb[p[off[11]]];
            d &= tb[p[off[5]]] | tb[p[off[13]]];
            d &= tb[p[off[7]]] | tb[p[off[15]]];
            if (!d) continue;
            unsigned bright = 0, dark = 0;
            for (int k = 0; k < 16; k++) {
                const int cls = tb[p[off[k]]];
                bright |= (unsigned)(cls >> 1) << k;
                dark |= (unsigned)(cls & 1) << k;
            }
            int corner = 0;
            for (int pol = 0; pol < 2 && !corner; pol++) {
                unsigned m = pol ? dark : bright, mm = m | (m << 16);
                unsigned a = mm & (mm >> 1);
                a &= a >> 2;
                a &= a >> 4;
                a &= mm >> 8;
                corner = (a & 0xffffu) != 0;
            }
            if (!corner) continue;
            is_corner[(size_t)y * w + x] = 1;
            score[(size_t)y * w + x] = nms ? (uint8_t)fast_score(p, off, threshold) : 0;
        }
    }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            size_t i = (size_t)y * w + x;
            if (!is_corner[i]) continue;
            int s = score[i];
            if (nms) {
                const uint8_t *r0 = score + i - w, *r1 = score + i, *r2 = score + i + w;
                if (!(s > r0[-1] && s > r0[0] && s > r0[1] && s > r1[-1] && s > r1[1] &&
                      s > r2[-1] && s > r2[0] && s > r2[1]))
                    continue;
            }
            if (n < cap) {
                orc_keypoint *k = &out[n];
                k->x = (float)x; k->y = (float)y; k->size = 7.f; k->angle = -1.f;
                k->response = (float)s; k->octave = 0; k->class_id = -1;
            }
            n++;
        }
    if (score != stackbuf) free(score);
    return n;
}

/* ======================================================================================= */
/* cv::GaussianBlur(7x7, sigma 2, REFLECT_101) for 8UC1 (S/ORBextractor.cc:1117)           */
/* ======================================================================================= */
void orc_gaussian_blur7(const uint8_t *src, int w, int h, int sstride,
                        uint8_t *dst, int dstride, int variant)
{
    static const int TAPS[2][7] = {{18, 34, 48, 56, 48, 34, 18}, {18, 34, 49, 55, 49, 34, 18}};
    const int *k = TAPS[variant ? 1 : 0];
    /* horizontal pass into 8.8 fixed point; each row is first copied into a buffer padded with its
     * REFLECT_101 continuation so the inner loop needs no border logic */
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    uint8_t *pad = (uint8_t *)malloc((size_t)w + 6);
    for (int y = 0; y < h; y++) {
        const uint8_t *s = src + (size_t)y * sstride;
        for (int i = 0; i < 3; i++) { pad[i] = s[reflect101(i - 3, w)]; pad[w + 3 + i] = s[reflect101(w + i, w)]; }
        memcpy(pad + 3, s, w);
        uint16_t *t = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            const uint8_t *q = pad + x;
            t[x] = (uint16_t)(k[0] * (q[0] + q[6]) + k[1] * (q[1] + q[5]) + k[2] * (q[2] + q[4]) + k[3] * q[3]);
        }
    }
    /* vertical pass; rows are picked through the same reflection; dst may alias src (tmp holds the input) */
    for (int y = 0; y < h; y++) {
        const uint16_t *r[7];
        for (int j = 0; j < 7; j++) r[j] = tmp + (size_t)reflect101(y + j - 3, h) * w;
        uint8_t *d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = (uint32_t)k[0] * (r[0][x] + r[6][x]) + (uint32_t)k[1] * (r[1][x] + r[5][x]) +
                           (uint32_t)k[2] * (r[2][x] + r[4][x]) + (uint32_t)k[3] * r[3][x];
            uint32_t v = (acc + 32768u) >> 16;      /* 16.16 -> u8, round half up, saturate */
            d[x] = (uint8_t)(v > 255 ? 255 : v);
        }
    }
    free(pad);
    free(tmp);
}

/* ======================================================================================= */
/* ORBextractor                                                                            */
/* ======================================================================================= */
#define PATCH_SIZE 31
#define HALF_PATCH 15
#define EDGE_THRESHOLD 19
#define MAX_LEVELS 32

static const int8_t PATTERN[1024] = {
#include "../include/orb_b200_pattern.inc"
};
const int8_t *orc_pattern(void) { return PATTERN; }

struct orc_extractor {
    int nfeatures, nlevels, ini_th, min_th, blur_variant;
    double scale_factor_d;  /* the header stores scaleFactor in a double (I/ORBextractor.h:98) */
    float scale[MAX_LEVELS], inv_scale[MAX_LEVELS], sigma2[MAX_LEVELS], inv_sigma2[MAX_LEVELS];
    int per_level[MAX_LEVELS];
    int umax[HALF_PATCH + 1];
    /* per-call stage state */
    int lw[MAX_LEVELS], lh[MAX_LEVELS];
    uint8_t *pix[MAX_LEVELS], *blur[MAX_LEVELS];
    orc_keypoint *cand[MAX_LEVELS], *kps[MAX_LEVELS];
    int ncand[MAX_LEVELS], nkps[MAX_LEVELS];
};

/* S/ORBextractor.cc:415-482 */
orc_extractor *orc_extractor_create(int nfeatures, float scaleFactor, int nlevels,
                                    int iniThFAST, int minThFAST)
{
    if (nlevels < 1 || nlevels > MAX_LEVELS) return NULL;
    orc_extractor *e = (orc_extractor *)calloc(1, sizeof(*e));
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->ini_th = iniThFAST; e->min_th = minThFAST;
    e->scale_factor_d = (double)scaleFactor;
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = (float)((double)e->scale[i - 1] * e->scale_factor_d);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        e->inv_scale[i] = 1.0f / e->scale[i];
        e->inv_sigma2[i] = 1.0f / e->sigma2[i];
    }
    /* geometric split of the feature budget (:444-455) */
    float factor = (float)(1.0 / e->scale_factor_d);
    /* evaluated in float except for the pow() itself (:445) */
    float desired = (float)nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->per_level[l] = orc_round(desired);
        sum += e->per_level[l];
        desired *= factor;
    }
    e->per_level[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    /* row extents of the circular orientation patch (:461-481) */
    int vmax = (int)floor(HALF_PATCH * sqrt(2.f) / 2 + 1);
    int vmin = (int)ceil(HALF_PATCH * sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (int v = 0; v <= vmax; ++v) e->umax[v] = (int)lrint(sqrt(hp2 - v * v));
    for (int v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

static void free_stage(orc_extractor *e)
{
    for (int l = 0; l < MAX_LEVELS; l++) {
        free(e->pix[l]); free(e->blur[l]); free(e->cand[l]); free(e->kps[l]);
        e->pix[l] = e->blur[l] = NULL; e->cand[l] = e->kps[l] = NULL;
        e->ncand[l] = e->nkps[l] = 0;
    }
}
void orc_extractor_destroy(orc_extractor *e) { if (e) { free_stage(e); free(e); } }
void orc_extractor_set_blur_variant(orc_extractor *e, int v) { e->blur_variant = v; }
const float *orc_scale_factors(const orc_extractor *e) { return e->scale; }
const float *orc_inv_scale_factors(const orc_extractor *e) { return e->inv_scale; }
const float *orc_level_sigma2(const orc_extractor *e) { return e->sigma2; }
const float *orc_inv_level_sigma2(const orc_extractor *e) { return e->inv_sigma2; }
const int *orc_features_per_level(const orc_extractor *e) { return e->per_level; }
const int *orc_umax(const orc_extractor *e) { return e->umax; }
int orc_level_width(const orc_extractor *e, int l) { return e->lw[l]; }
int orc_level_height(const orc_extractor *e, int l) { return e->lh[l]; }
const uint8_t *orc_level_pixels(const orc_extractor *e, int l) { return e->pix[l]; }
const uint8_t *orc_level_blurred(const orc_extractor *e, int l) { return e->blur[l]; }
int orc_level_candidates(const orc_extractor *e, int l, const orc_keypoint **p) { *p = e->cand[l]; return e->ncand[l]; }
int orc_level_keypoints(const orc_extractor *e, int l, const orc_keypoint **p) { *p = e->kps[l]; return e->nkps[l]; }

/* --------------------------------------------------------------------------------------- */
/* DistributeOctTree (S/ORBextractor.cc:552-776) with DivideNode (:494-550).               */
/* Sequential restatement on a doubly linked list held in arrays.  Node memory is never    */
/* reused, so a node's index is its creation sequence number: that is the order the         */
/* reference's pointer-valued tie-break (:694-698) takes under a monotonic allocator.       */
/* --------------------------------------------------------------------------------------- */
typedef struct {
    int x0, x1, y0, y1;      /* UL.x, UR.x, UL.y, BL.y */
    int first, count;        /* slice of the point-index pool */
    int prev, next;          /* list links, -1 = none */
    int no_more;
} qnode;

typedef struct {
    qnode *nodes; int nnodes, capnodes;
    int *pool; int npool, cappool;
    int head, tail, size;
} qtree;

static int q_new_node(qtree *t)
{
    if (t->nnodes == t->capnodes) {
        t->capnodes *= 2;
        t->nodes = (qnode *)realloc(t->nodes, sizeof(qnode) * t->capnodes);
    }
    qnode *n = &t->nodes[t->nnodes];
    memset(n, 0, sizeof(*n));
    n->prev = n->next = -1;
    return t->nnodes++;
}
static int q_pool_reserve(qtree *t, int n)
{
    while (t->npool + n > t->cappool) {
        t->cappool *= 2;
        t->pool = (int *)realloc(t->pool, sizeof(int) * t->cappool);
    }
    int at = t->npool;
    t->npool += n;
    return at;
}
static void q_push_front(qtree *t, int id)
{
    t->nodes[id].prev = -1; t->nodes[id].next = t->head;
    if (t->head >= 0) t->nodes[t->head].prev = id; else t->tail = id;
    t->head = id; t->size++;
}
static void q_push_back(qtree *t, int id)
{
    t->nodes[id].next = -1; t->nodes[id].prev = t->tail;
    if (t->tail >= 0) t->nodes[t->tail].next = id; else t->head = id;
    t->tail = id; t->size++;
}
static int q_erase(qtree *t, int id) /* returns the following node */
{
    int p = t->nodes[id].prev, n = t->nodes[id].next;
    if (p >= 0) t->nodes[p].next = n; else t->head = n;
    if (n >= 0) t->nodes[n].prev = p; else t->tail = p;
    t->size--;
    return n;
}

/* Split node `id` into up to four children pushed to the list front in n1..n4 order;
 * children holding more than one point are appended to `expand` (id list). */
static void q_divide(qtree *t, const orc_keypoint *pts, int id, int *expand, int *nexpand)
{
    qnode P = t->nodes[id];
    int halfX = (int)ceilf((float)(P.x1 - P.x0) / 2), halfY = (int)ceilf((float)(P.y1 - P.y0) / 2);
    int midx = P.x0 + halfX, midy = P.y0 + halfY;
    int cx0[4] = {P.x0, midx, P.x0, midx}, cx1[4] = {midx, P.x1, midx, P.x1};
    int cy0[4] = {P.y0, P.y0, midy, midy}, cy1[4] = {midy, midy, P.y1, P.y1};
    int cnt[4] = {0, 0, 0, 0};
    for (int i = 0; i < P.count; i++) {
        const orc_keypoint *k = &pts[t->pool[P.first + i]];
        int q = (k->x < (float)midx ? 0 : 1) + (k->y < (float)midy ? 0 : 2);
        cnt[q]++;
    }
    int at[4], base = q_pool_reserve(t, P.count);
    at[0] = base; at[1] = at[0] + cnt[0]; at[2] = at[1] + cnt[1]; at[3] = at[2] + cnt[2];
    int fill[4] = {0, 0, 0, 0};
    for (int i = 0; i < P.count; i++) {
        int pi = t->pool[P.first + i];
        const orc_keypoint *k = &pts[pi];
        int q = (k->x < (float)midx ? 0 : 1) + (k->y < (float)midy ? 0 : 2);
        t->pool[at[q] + fill[q]++] = pi;
    }
    for (int q = 0; q < 4; q++) {
        if (cnt[q] == 0) continue;
        int c = q_new_node(t);
        qnode *n = &t->nodes[c];
        n->x0 = cx0[q]; n->x1 = cx1[q]; n->y0 = cy0[q]; n->y1 = cy1[q];
        n->first = at[q]; n->count = cnt[q]; n->no_more = cnt[q] == 1;
        q_push_front(t, c);
        if (cnt[q] > 1) expand[(*nexpand)++] = c;
    }
}

typedef struct { int count, id; } qkey;
static int qkey_cmp_asc(const void *a, const void *b)
{
    const qkey *x = (const qkey *)a, *y = (const qkey *)b;
    if (x->count != y->count) return x->count < y->count ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id ? 1 : 0);
}
static int qkey_cmp_asc_rev(const void *a, const void *b)
{
    const qkey *x = (const qkey *)a, *y = (const qkey *)b;
    if (x->count != y->count) return x->count < y->count ? -1 : 1;
    return x->id > y->id ? -1 : (x->id < y->id ? 1 : 0);
}

int orc_distribute_octree(const orc_keypoint *in, int n, int minX, int maxX, int minY, int maxY,
                          int N, int tie_break, orc_keypoint *out, int cap)
{
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return -2;
    const float hX = (float)(maxX - minX) / nIni;
    qtree t;
    t.capnodes = 64 + 8 * (n > N ? N : n) + 4 * nIni; t.nnodes = 0;
    t.nodes = (qnode *)malloc(sizeof(qnode) * t.capnodes);
    t.cappool = 16 * (n + 16); t.npool = 0;
    t.pool = (int *)malloc(sizeof(int) * t.cappool);
    t.head = t.tail = -1; t.size = 0;

    /* root nodes (:564-576) and point assignment (:579-583) */
    int *rootcnt = (int *)calloc(nIni, sizeof(int)), *rootof = (int *)malloc(sizeof(int) * (n + 1));
    for (int i = 0; i < n; i++) {
        int r = (int)(in[i].x / hX);
        if (r < 0) r = 0;
        if (r >= nIni) r = nIni - 1;   /* cannot happen for in-range points */
        rootof[i] = r; rootcnt[r]++;
    }
    for (int i = 0; i < nIni; i++) {
        int id = q_new_node(&t);
        qnode *nd = &t.nodes[id];
        nd->x0 = (int)(hX * (float)i); nd->x1 = (int)(hX * (float)(i + 1));
        nd->y0 = 0; nd->y1 = maxY - minY;
        nd->first = q_pool_reserve(&t, rootcnt[i]); nd->count = 0;
        q_push_back(&t, id);
    }
    for (int i = 0; i < n; i++) {
        qnode *nd = &t.nodes[rootof[i]];
        t.pool[nd->first + nd->count++] = i;
    }
    for (int id = t.head; id >= 0;) {       /* :585-598 */
        qnode *nd = &t.nodes[id];
        if (nd->count == 1) { nd->no_more = 1; id = nd->next; }
        else if (nd->count == 0) id = q_erase(&t, id);
        else id = nd->next;
    }

    /* a node is expandable only if it holds >= 2 points and live nodes are disjoint, so at
     * most n/2 expandable nodes exist at any time */
    int capexp = n / 2 + 8;
    int *expand = (int *)malloc(sizeof(int) * capexp), nexpand = 0;
    qkey *order = (qkey *)malloc(sizeof(qkey) * capexp);
    int finish = 0;
    while (!finish) {
        int prev_size = t.size;
        nexpand = 0;
        /* breadth pass: split every expandable node, front to back (:613-678) */
        for (int id = t.head; id >= 0;) {
            if (t.nodes[id].no_more) { id = t.nodes[id].next; continue; }
            q_divide(&t, in, id, expand, &nexpand);
            id = q_erase(&t, id);
        }
        if (t.size >= N || t.size == prev_size) {
            finish = 1;
        } else if (t.size + nexpand * 3 > N) {
            /* largest-first splitting until N leaves (:686-751) */
            while (!finish) {
                prev_size = t.size;
                int nprev = nexpand;
                for (int i = 0; i < nprev; i++) { order[i].count = t.nodes[expand[i]].count; order[i].id = expand[i]; }
                qsort(order, nprev, sizeof(qkey), tie_break ? qkey_cmp_asc_rev : qkey_cmp_asc);
                nexpand = 0;
                for (int j = nprev - 1; j >= 0; j--) {
                    q_divide(&t, in, order[j].id, expand, &nexpand);
                    q_erase(&t, order[j].id);
                    if (t.size >= N) break;
                }
                if (t.size >= N || t.size == prev_size) finish = 1;
            }
        }
    }

    /* best response per leaf, first point wins ties (:755-773) */
    int nout = 0;
    for (int id = t.head; id >= 0; id = t.nodes[id].next) {
        const qnode *nd = &t.nodes[id];
        int best = t.pool[nd->first];
        for (int k = 1; k < nd->count; k++) {
            int pi = t.pool[nd->first + k];
            if (in[pi].response > in[best].response) best = pi;
        }
        if (nout < cap) out[nout] = in[best];
        nout++;
    }
    free(order); free(expand); free(rootof); free(rootcnt); free(t.pool); free(t.nodes);
    return nout <= cap ? nout : -1;
}

/* --------------------------------------------------------------------------------------- */
/* IC_Angle (S/ORBextractor.cc:82-109)                                                     */
/* --------------------------------------------------------------------------------------- */
static float ic_angle(const uint8_t *img, int stride, int px, int py, const int *umax)
{
    const uint8_t *c = img + (size_t)py * stride + px;
    int m01 = 0, m10 = 0;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m10 += u * c[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int vs = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int a = c[u + v * stride], b = c[u - v * stride];
            vs += a - b;
            m10 += u * (a + b);
        }
        m01 += v * vs;
    }
    return orc_fast_atan2((float)m01, (float)m10);
}

/* computeOrbDescriptor (S/ORBextractor.cc:113-152): every product and sum is a separately
 * rounded fp32 operation (x86-64 SSE, no contraction), cvRound is round-half-even. */
static void orb_descriptor(const uint8_t *img, int stride, int px, int py, float angle_deg, uint8_t *desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factorPI;
    float a, b;
    orc_sincosf(angle, &b, &a);
    const uint8_t *c = img + (size_t)py * stride + px;
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int t[2];
            for (int s = 0; s < 2; s++) {
                const int8_t *pp = &PATTERN[((i * 16) + 2 * k + s) * 2];
                volatile float xb = (float)pp[0] * b, ya = (float)pp[1] * a;
                volatile float xa = (float)pp[0] * a, yb = (float)pp[1] * b;
                volatile float fy = xb + ya, fx = xa - yb;
                t[s] = c[orc_round(fy) * stride + orc_round(fx)];
            }
            val |= (t[0] < t[1]) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

/* --------------------------------------------------------------------------------------- */
/* operator() (S/ORBextractor.cc:1064-1136): ComputePyramid (:1138-1168),                  */
/* ComputeKeyPointsOctTree (:778-873), blur + descriptors + rescale                        */
/* --------------------------------------------------------------------------------------- */
int orc_extract(orc_extractor *e, const uint8_t *img, int w, int h, int stride,
                orc_keypoint *kps, uint8_t *desc, int cap)
{
    free_stage(e);
    if (!img || w <= 0 || h <= 0) return 0;               /* empty image: nothing happens (:1068) */

    /* pyramid: each level is resized from the previous level (:1157) */
    for (int l = 0; l < e->nlevels; l++) {
        e->lw[l] = orc_round((float)w * e->inv_scale[l]);
        e->lh[l] = orc_round((float)h * e->inv_scale[l]);
        if (e->lw[l] - 2 * 16 < 30 || e->lh[l] - 2 * 16 < 30) return -2;
    }
    for (int l = 0; l < e->nlevels; l++) {
        e->pix[l] = (uint8_t *)malloc((size_t)e->lw[l] * e->lh[l]);
        if (l == 0)
            for (int y = 0; y < h; y++) memcpy(e->pix[0] + (size_t)y * w, img + (size_t)y * stride, w);
        else
            orc_resize_linear_u8(e->pix[l - 1], e->lw[l - 1], e->lh[l - 1], e->lw[l - 1],
                                 e->pix[l], e->lw[l], e->lh[l], e->lw[l]);
    }

    const float W = 30;
    for (int l = 0; l < e->nlevels; l++) {
        const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
        const int maxBX = e->lw[l] - EDGE_THRESHOLD + 3, maxBY = e->lh[l] - EDGE_THRESHOLD + 3;
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
        if ((int)roundf((float)(maxBX - minBX) / (maxBY - minBY)) < 1) return -2;

        int capc = ((maxBX - minBX) / 2 + 2) * ((maxBY - minBY) / 2 + 2);
        orc_keypoint *cand = (orc_keypoint *)malloc(sizeof(orc_keypoint) * capc);
        orc_keypoint *cell = (orc_keypoint *)malloc(sizeof(orc_keypoint) * 64 * 64);
        int nc = 0;
        for (int i = 0; i < nRows; i++) {
            const int iniY = minBY + i * hCell;
            int maxY = iniY + hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = maxBY;
            for (int j = 0; j < nCols; j++) {
                const int iniX = minBX + j * wCell;
                int maxX = iniX + wCell + 6;
                if (iniX >= maxBX - 6) continue;
                if (maxX > maxBX) maxX = maxBX;
                const uint8_t *win = e->pix[l] + (size_t)iniY * e->lw[l] + iniX;
                int n = orc_fast9_16(win, maxX - iniX, maxY - iniY, e->lw[l], e->ini_th, 1, cell, 64 * 64);
                if (n == 0)
                    n = orc_fast9_16(win, maxX - iniX, maxY - iniY, e->lw[l], e->min_th, 1, cell, 64 * 64);
                for (int k = 0; k < n; k++) {
                    cell[k].x += j * wCell;
                    cell[k].y += i * hCell;
                    cand[nc++] = cell[k];
                }
            }
        }
        free(cell);
        e->cand[l] = cand; e->ncand[l] = nc;

        int capk = e->per_level[l] + 4 * 64 + 8;
        orc_keypoint *lk = (orc_keypoint *)malloc(sizeof(orc_keypoint) * capk);
        int nk = orc_distribute_octree(cand, nc, minBX, maxBX, minBY, maxBY, e->per_level[l], 0, lk, capk);
        if (nk < 0) { free(lk); return -2; }
        const int scaledPatchSize = (int)(PATCH_SIZE * e->scale[l]);
        for (int k = 0; k < nk; k++) {
            lk[k].x += minBX; lk[k].y += minBY;
            lk[k].octave = l; lk[k].size = (float)scaledPatchSize;
        }
        e->kps[l] = lk; e->nkps[l] = nk;
    }
    for (int l = 0; l < e->nlevels; l++)
        for (int k = 0; k < e->nkps[l]; k++) {
            orc_keypoint *kp = &e->kps[l][k];
            kp->angle = ic_angle(e->pix[l], e->lw[l], orc_round(kp->x), orc_round(kp->y), e->umax);
        }

    int total = 0;
    for (int l = 0; l < e->nlevels; l++) total += e->nkps[l];
    if (total > cap) return -1;
    int off = 0;
    for (int l = 0; l < e->nlevels; l++) {
        if (e->nkps[l] == 0) continue;
        e->blur[l] = (uint8_t *)malloc((size_t)e->lw[l] * e->lh[l]);
        orc_gaussian_blur7(e->pix[l], e->lw[l], e->lh[l], e->lw[l], e->blur[l], e->lw[l], e->blur_variant);
        for (int k = 0; k < e->nkps[l]; k++) {
            const orc_keypoint *kp = &e->kps[l][k];
            orb_descriptor(e->blur[l], e->lw[l], orc_round(kp->x), orc_round(kp->y), kp->angle,
                           desc + (size_t)(off + k) * 32);
            kps[off + k] = *kp;
            if (l != 0) {
                kps[off + k].x = kp->x * e->scale[l];
                kps[off + k].y = kp->y * e->scale[l];
            }
        }
        off += e->nkps[l];
    }
    return total;
}
