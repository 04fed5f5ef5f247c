/* orbx CPU oracle, extractor part — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * Restates ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc) and the OpenCV primitives it
 * delegates to, in dependency-free C.  Compile with -ffp-contract=off: every float expression below
 * is meant to round after each operation, like the reference built without FMA contraction.
 *
 * Written-down choices where the reference is build- or heap-dependent (SURVEY.md §7 hard parts):
 *  - DistributeOctTree sorts (size, node pointer) pairs (ORBextractor.cc:684).  Pointer order is heap
 *    layout; we define it as CREATION ORDER (a bump allocator): equal sizes => later-created node is
 *    "larger" and therefore expanded first.
 *  - cos/sin of the keypoint angle (ORBextractor.cc:112-113, std::cos(float) == cosf): we use a
 *    double-precision Cody-Waite + polynomial evaluation rounded once to float (orbo_sincos_f);
 *    tests/test_oracle_cv2.py checks it equals glibc cosf/sinf on millions of angles.  The CUDA path
 *    repeats the same operation sequence, so GPU == oracle holds by construction.
 */
#include "orbx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <float.h>
#include <stddef.h>

#define EDGE_THRESHOLD 19
#define HALF_PATCH 15
#define PATCH_SIZE 31
#define MAX_LEVELS 16

static const int8_t k_pattern[1024] = {
#include "orb_pattern.inc"
};

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* cvRound(float): SSE cvtss2si == round-half-to-even under the default rounding mode */
int orbo_cv_round_f(float v) { return (int)lrintf(v); }
static int cv_round_d(double v) { return (int)lrint(v); }

/* ------------------------------------------------------------------------------------------------
 * cv::resize(8UC1, INTER_LINEAR)  (OpenCV imgproc/resize.cpp: resizeGeneric_ + HResizeLinear +
 * VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>>; INTER_RESIZE_COEF_BITS = 11).
 * Called at ORBextractor.cc:1120.
 * ---------------------------------------------------------------------------------------------- */
void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh,
                           int dstride) {
    const double inv_sx = (double)dw / sw, inv_sy = (double)dh / sh;
    const double scale_x = 1. / inv_sx, scale_y = 1. / inv_sy;
    int *xofs = (int *)malloc(sizeof(int) * dw);
    short *ia = (short *)malloc(sizeof(short) * 2 * dw);
    int *row0 = (int *)malloc(sizeof(int) * dw), *row1 = (int *)malloc(sizeof(int) * dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ia[2 * dx] = (short)orbo_cv_round_f((1.f - fx) * 2048);
        ia[2 * dx + 1] = (short)orbo_cv_round_f(fx * 2048);
    }
    int prev_sy0 = -2, prev_sy1 = -2;
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        short b0 = (short)orbo_cv_round_f((1.f - fy) * 2048), b1 = (short)orbo_cv_round_f(fy * 2048);
        int sy0 = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);
        int sy1 = sy + 1 < 0 ? 0 : (sy + 1 >= sh ? sh - 1 : sy + 1);
        /* horizontal pass of the two source rows (cache the previous pair like the row ring of OpenCV) */
        if (sy0 == prev_sy1) {
            int *t = row0; row0 = row1; row1 = t;
        } else if (sy0 != prev_sy0) {
            const uint8_t *S = src + (size_t)sy0 * sstride;
            for (int dx = 0; dx < dw; dx++) {
                int sx = xofs[dx];
                int s1 = sx + 1 < sw ? S[sx + 1] : S[sx];
                row0[dx] = S[sx] * ia[2 * dx] + s1 * ia[2 * dx + 1];
            }
        }
        if (sy1 == sy0) {
            memcpy(row1, row0, sizeof(int) * dw);
        } else if (!(sy0 == prev_sy0 && sy1 == prev_sy1)) {
            const uint8_t *S = src + (size_t)sy1 * sstride;
            for (int dx = 0; dx < dw; dx++) {
                int sx = xofs[dx];
                int s1 = sx + 1 < sw ? S[sx + 1] : S[sx];
                row1[dx] = S[sx] * ia[2 * dx] + s1 * ia[2 * dx + 1];
            }
        }
        prev_sy0 = sy0; prev_sy1 = sy1;
        uint8_t *D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++) {
            int v = (((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs); free(ia); free(row0); free(row1);
}

/* cv::copyMakeBorder(..., BORDER_REFLECT_101) around an interior already in place
 * (ORBextractor.cc:1122-1128).  buf points at the interior origin; pad pixels exist on all sides. */
void orbo_border_reflect101(uint8_t *buf, int w, int h, int stride, int pad) {
    for (int y = 0; y < h; y++) {
        uint8_t *r = buf + (size_t)y * stride;
        for (int i = 1; i <= pad; i++) {
            r[-i] = r[i];
            r[w - 1 + i] = r[w - 1 - i];
        }
    }
    for (int i = 1; i <= pad; i++) {
        memcpy(buf + (ptrdiff_t)(-i) * stride - pad, buf + (ptrdiff_t)i * stride - pad, w + 2 * pad);
        memcpy(buf + (ptrdiff_t)(h - 1 + i) * stride - pad, buf + (ptrdiff_t)(h - 1 - i) * stride - pad, w + 2 * pad);
    }
}

/* cv::GaussianBlur(8U, Size(7,7), 2, 2, BORDER_REFLECT_101) (ORBextractor.cc:1086): OpenCV's
 * fixed-point path, Q8.8 taps {18,34,48,56,48,34,18}, rows then columns, (acc + 2^15) >> 16. */
static const int k_g7[7] = {18, 34, 48, 56, 48, 34, 18};
static inline int refl101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}
void orbo_gaussian7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *S = src + (size_t)y * sstride;
        uint16_t *T = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            int acc = 0;
            if (x >= 3 && x < w - 3)
                for (int k = 0; k < 7; k++) acc += k_g7[k] * S[x + k - 3];
            else
                for (int k = 0; k < 7; k++) acc += k_g7[k] * S[refl101(x + k - 3, w)];
            T[x] = (uint16_t)acc;
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *R[7];
        for (int k = 0; k < 7; k++) R[k] = tmp + (size_t)refl101(y + k - 3, h) * w;
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; k++) acc += (uint32_t)k_g7[k] * R[k][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------------
 * cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16 (OpenCV features2d/fast.cpp +
 * fast_score.cpp).  Called per 30-px cell at ORBextractor.cc:809/:814.
 * score = (max over the 16 arcs of 9 contiguous ring pixels of min|v-p| with one sign) - 1;
 * corner iff score >= threshold; NMS: strictly greater than the 8 neighbours, where pixels that are
 * not corners or lie outside x in [3,w-3), y in [3,h-3) count as 0.  Output raster order.
 * ---------------------------------------------------------------------------------------------- */
static const int k_ring_dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int k_ring_dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

static inline int fast_score_at(const uint8_t *p, const int off[16], int threshold) {
    const int v = p[0];
    int d[25];
    /* quick reject like OpenCV's cascade (fast.cpp: d = tab[0] | tab[8]; d &= tab[2] | tab[10]; ...): a 9-arc covers at least one
     * pixel of every opposite pair (k, k + 8), with the arc's polarity.  bit 0 = darker ring pixel (v - p > t), bit 1 = brighter.
     * Only a necessary condition; the arc scan below decides. */
    int m = 3;
    for (int k = 0; k < 8 && m; k += 2) {          /* even pairs first, like OpenCV */
        const int a = v - p[off[k]], b = v - p[off[k + 8]];
        m &= ((a > threshold) | ((a < -threshold) << 1)) | ((b > threshold) | ((b < -threshold) << 1));
    }
    for (int k = 1; k < 8 && m; k += 2) {
        const int a = v - p[off[k]], b = v - p[off[k + 8]];
        m &= ((a > threshold) | ((a < -threshold) << 1)) | ((b > threshold) | ((b < -threshold) << 1));
    }
    if (!m) return 0;
    for (int k = 0; k < 16; k++) d[k] = v - p[off[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = 0; /* max over arcs of min(d) (darker ring) and of min(-d) (brighter ring) */
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) {
            int t = d[k + j];
            if (t < mn) mn = t;
            if (t > mx) mx = t;
        }
        if (mn > best) best = mn;
        if (-mx > best) best = -mx;
    }
    /* corner iff 9 contiguous differ by more than threshold  <=>  best > threshold */
    return best > threshold ? best - 1 : 0;
}

/* scores: w*h uint8 map (0 where not a corner / outside the valid ring) */
static void fast_score_map(const uint8_t *img, int w, int h, int stride, int threshold, uint8_t *scores) {
    int off[16];
    for (int k = 0; k < 16; k++) off[k] = k_ring_dy[k] * stride + k_ring_dx[k];
    memset(scores, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *row = img + (size_t)y * stride;
        uint8_t *srow = scores + (size_t)y * w;
        for (int x = 3; x < w - 3; x++) srow[x] = (uint8_t)fast_score_at(row + x, off, threshold);
    }
}

int orbo_fast9(const uint8_t *img, int w, int h, int stride, int threshold, int *xs, int *ys, int *scores,
               int cap) {
    if (w < 7 || h < 7) return 0;
    uint8_t *sm = (uint8_t *)malloc((size_t)w * h);
    fast_score_map(img, w, h, stride, threshold, sm);
    int n = 0;
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int s = sm[y * w + x];
            if (!s) continue;
            const uint8_t *c = sm + y * w + x;
            if (s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] && s > c[w] &&
                s > c[w + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; scores[n] = s; }
                n++;
            }
        }
    free(sm);
    return n;
}

/* cv::fastAtan2 (OpenCV core/mathfuncs_core: atan_f32), degrees; called at ORBextractor.cc:103 */
float orbo_fast_atan2(float y, float x) {
    const float sc = (float)(180 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc;
    const float p5 = 0.1555786518463281f * sc, p7 = -0.04432655554792128f * sc;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* sinf/cosf stand-in: double Cody-Waite reduction by pi/2 + degree-13/12 polynomials, rounded once
 * to float.  Sequence of IEEE double mul/add only (no FMA), so the CUDA kernel can repeat it exactly. */
void orbo_sincos_f(float xf, float *s, float *c) {
    const double x = (double)xf;
    const double kf = rint(x * 6.36619772367581382433e-01);
    double r = x - kf * 1.57079632673412561417e+00;
    r = r - kf * 6.07710050650619224932e-11;
    const double z = r * r;
    double ps = 1.58969099521155010221e-10;
    ps = ps * z + -2.50507602534068634195e-08;
    ps = ps * z + 2.75573137070700676789e-06;
    ps = ps * z + -1.98412698298579493134e-04;
    ps = ps * z + 8.33333333332248946124e-03;
    ps = ps * z + -1.66666666666666324348e-01;
    const double sn = r + (r * z) * ps;
    double pc = -1.13596475577881948265e-11;
    pc = pc * z + 2.08757232129817482790e-09;
    pc = pc * z + -2.75573143513906633035e-07;
    pc = pc * z + 2.48015872894767294178e-05;
    pc = pc * z + -1.38888888888741095749e-03;
    pc = pc * z + 4.16666666666666019037e-02;
    const double cs = (1.0 - 0.5 * z) + (z * z) * pc;
    const long k = (long)kf & 3;
    double so, co;
    switch (k) {
    case 0: so = sn; co = cs; break;
    case 1: so = cs; co = -sn; break;
    case 2: so = -sn; co = -cs; break;
    default: so = -cs; co = sn; break;
    }
    *s = (float)so;
    *c = (float)co;
}

/* ------------------------------------------------------------------------------------------------ */
struct orbo_extractor {
    int nfeatures, nlevels, ini_th, min_th;
    double scale_factor; /* ORBextractor.h: `double scaleFactor` initialised from a float */
    float scale[MAX_LEVELS], inv_scale[MAX_LEVELS], sigma2[MAX_LEVELS], inv_sigma2[MAX_LEVELS];
    int quota[MAX_LEVELS];
    int umax[HALF_PATCH + 1];
    /* per-call state */
    uint8_t *pyr[MAX_LEVELS]; /* padded buffers */
    int lw[MAX_LEVELS], lh[MAX_LEVELS], lstride[MAX_LEVELS];
    size_t pyr_cap[MAX_LEVELS];
    orbo_keypoint *cand[MAX_LEVELS];
    int ncand[MAX_LEVELS], cand_cap[MAX_LEVELS];
    double t_stage[6];
};

/* ORBextractor::ORBextractor, ORBextractor.cc:410-470 */
orbo_extractor *orbo_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    if (nlevels < 1 || nlevels > MAX_LEVELS) return NULL;
    orbo_extractor *e = (orbo_extractor *)calloc(1, sizeof(*e));
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->ini_th = ini_th; e->min_th = min_th;
    e->scale_factor = (double)scale_factor;
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scale_factor); /* float*double -> float, :421 */
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        e->inv_scale[i] = 1.0f / e->scale[i];
        e->inv_sigma2[i] = 1.0f / e->sigma2[i];
    }
    float factor = (float)(1.0f / e->scale_factor);
    float n_desired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->quota[l] = orbo_cv_round_f(n_desired);
        sum += e->quota[l];
        n_desired *= factor;
    }
    e->quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    /* umax, :454-469 */
    int v, v0, vmax = (int)floorf(HALF_PATCH * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(HALF_PATCH * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH * HALF_PATCH;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

void orbo_extractor_destroy(orbo_extractor *e) {
    if (!e) return;
    for (int l = 0; l < MAX_LEVELS; l++) { free(e->pyr[l]); free(e->cand[l]); }
    free(e);
}

void orbo_extractor_tables(const orbo_extractor *e, float *scale, float *inv_scale, float *sigma2,
                           float *inv_sigma2, int *quota, int *umax16) {
    for (int l = 0; l < e->nlevels; l++) {
        if (scale) scale[l] = e->scale[l];
        if (inv_scale) inv_scale[l] = e->inv_scale[l];
        if (sigma2) sigma2[l] = e->sigma2[l];
        if (inv_sigma2) inv_sigma2[l] = e->inv_sigma2[l];
        if (quota) quota[l] = e->quota[l];
    }
    if (umax16) for (int v = 0; v <= HALF_PATCH; v++) umax16[v] = e->umax[v];
}

/* DistributeOctTree stops once the node list holds >= quota nodes; one DivideNode adds at most 3.  The
 * first pass always runs, so a level can also return up to 4*nIni nodes even when its quota is smaller;
 * nIni <= 16 is enforced in orbo_extract. */
#define NINI_MAX 16
int orbo_extractor_capacity(const orbo_extractor *e) {
    int c = 0;
    for (int l = 0; l < e->nlevels; l++) c += e->quota[l] + 3 > 4 * NINI_MAX ? e->quota[l] + 3 : 4 * NINI_MAX;
    return c;
}

/* ORBextractor::ComputePyramid, :1107-1132 */
static int compute_pyramid(orbo_extractor *e, const uint8_t *img, int w, int h, int stride) {
    for (int l = 0; l < e->nlevels; l++) {
        float sc = e->inv_scale[l];
        int lw = orbo_cv_round_f((float)w * sc), lh = orbo_cv_round_f((float)h * sc);
        if (lw <= 2 * EDGE_THRESHOLD || lh <= 2 * EDGE_THRESHOLD) return -1; /* reflect101 needs pad < size */
        int st = lw + 2 * EDGE_THRESHOLD;
        size_t need = (size_t)st * (lh + 2 * EDGE_THRESHOLD);
        if (need > e->pyr_cap[l]) {
            free(e->pyr[l]);
            e->pyr[l] = (uint8_t *)malloc(need);
            e->pyr_cap[l] = need;
        }
        e->lw[l] = lw; e->lh[l] = lh; e->lstride[l] = st;
        uint8_t *in = e->pyr[l] + (size_t)EDGE_THRESHOLD * st + EDGE_THRESHOLD;
        if (l == 0) {
            for (int y = 0; y < h; y++) memcpy(in + (size_t)y * st, img + (size_t)y * stride, w);
        } else {
            const uint8_t *pin = e->pyr[l - 1] + (size_t)EDGE_THRESHOLD * e->lstride[l - 1] + EDGE_THRESHOLD;
            orbo_resize_linear_u8(pin, e->lw[l - 1], e->lh[l - 1], e->lstride[l - 1], in, lw, lh, st);
        }
        orbo_border_reflect101(in, lw, lh, st, EDGE_THRESHOLD);
    }
    return 0;
}

static void cand_push(orbo_extractor *e, int l, float x, float y, float resp) {
    if (e->ncand[l] == e->cand_cap[l]) {
        e->cand_cap[l] = e->cand_cap[l] ? 2 * e->cand_cap[l] : 4096;
        e->cand[l] = (orbo_keypoint *)realloc(e->cand[l], sizeof(orbo_keypoint) * e->cand_cap[l]);
    }
    orbo_keypoint *k = &e->cand[l][e->ncand[l]++];
    k->x = x; k->y = y; k->size = 7.f; k->angle = -1.f; k->response = resp; k->octave = 0; k->class_id = -1;
}

/* per-cell FAST of ComputeKeyPointsOctTree, :765-829.  Coordinates are relative to (minBorder, minBorder). */
static void fast_cells(orbo_extractor *e, int l) {
    const float W = 30;
    const uint8_t *in = e->pyr[l] + (size_t)EDGE_THRESHOLD * e->lstride[l] + EDGE_THRESHOLD;
    const int st = e->lstride[l];
    const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
    const int maxBX = e->lw[l] - EDGE_THRESHOLD + 3, maxBY = e->lh[l] - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    e->ncand[l] = 0;
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)ceilf(width / nCols), hCell = (int)ceilf(height / nRows);
    int xs[4096], ys[4096], sc[4096];
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
            const uint8_t *cell = in + (size_t)y0 * st + x0;
            int n = orbo_fast9(cell, cw, ch, st, e->ini_th, xs, ys, sc, 4096);
            if (n == 0) n = orbo_fast9(cell, cw, ch, st, e->min_th, xs, ys, sc, 4096);
            for (int k = 0; k < n; k++)
                cand_push(e, l, (float)xs[k] + j * wCell, (float)ys[k] + i * hCell, (float)sc[k]);
        }
    }
}

/* ---- DistributeOctTree, :539-763, with ExtractorNode::DivideNode :481-537 ---- */
typedef struct {
    int x0, x1, y0, y1;
    int *keys;
    int nkeys;
    int no_more;
    int prev, next; /* list links */
    long seq;       /* creation order: stands in for the node's address */
} onode;

typedef struct {
    onode *n;
    int count, cap;
    int head, tail, size;
    long seq;
} olist;

static int ol_new(olist *L) {
    if (L->count == L->cap) {
        L->cap = L->cap ? L->cap * 2 : 256;
        L->n = (onode *)realloc(L->n, sizeof(onode) * L->cap);
    }
    onode *nd = &L->n[L->count];
    memset(nd, 0, sizeof(*nd));
    nd->prev = nd->next = -1;
    nd->seq = L->seq++;
    return L->count++;
}
static void ol_push_back(olist *L, int id) {
    L->n[id].prev = L->tail; L->n[id].next = -1;
    if (L->tail >= 0) L->n[L->tail].next = id; else L->head = id;
    L->tail = id; L->size++;
}
static void ol_push_front(olist *L, int id) {
    L->n[id].next = L->head; L->n[id].prev = -1;
    if (L->head >= 0) L->n[L->head].prev = id; else L->tail = id;
    L->head = id; L->size++;
}
static int ol_erase(olist *L, int id) { /* returns next */
    int p = L->n[id].prev, nx = L->n[id].next;
    if (p >= 0) L->n[p].next = nx; else L->head = nx;
    if (nx >= 0) L->n[nx].prev = p; else L->tail = p;
    L->size--;
    free(L->n[id].keys); L->n[id].keys = NULL;
    return nx;
}

/* divides node id; children ids (or -1 if empty) in c[4]; child nodes are NOT yet linked */
static void divide_node(olist *L, int id, const orbo_keypoint *K, int c[4]) {
    const int x0 = L->n[id].x0, x1 = L->n[id].x1, y0 = L->n[id].y0, y1 = L->n[id].y1;
    const int halfX = (int)ceilf((float)(x1 - x0) / 2), halfY = (int)ceilf((float)(y1 - y0) / 2);
    const int nk = L->n[id].nkeys;
    int *buf[4], cnt[4] = {0, 0, 0, 0};
    for (int q = 0; q < 4; q++) buf[q] = (int *)malloc(sizeof(int) * (nk ? nk : 1));
    const int mx = x0 + halfX, my = y0 + halfY;
    for (int i = 0; i < nk; i++) {
        const int ki = L->n[id].keys[i];
        const orbo_keypoint *kp = &K[ki];
        int q;
        if (kp->x < (float)mx) q = (kp->y < (float)my) ? 0 : 2;
        else q = (kp->y < (float)my) ? 1 : 3;
        buf[q][cnt[q]++] = ki;
    }
    const int bx0[4] = {x0, mx, x0, mx}, bx1[4] = {mx, x1, mx, x1};
    const int by0[4] = {y0, y0, my, my}, by1[4] = {my, my, y1, y1};
    for (int q = 0; q < 4; q++) {
        if (cnt[q] == 0) { free(buf[q]); c[q] = -1; continue; }
        int ch = ol_new(L);
        onode *nd = &L->n[ch];
        nd->x0 = bx0[q]; nd->x1 = bx1[q]; nd->y0 = by0[q]; nd->y1 = by1[q];
        nd->keys = buf[q]; nd->nkeys = cnt[q];
        nd->no_more = (cnt[q] == 1);
        c[q] = ch;
    }
}

typedef struct { int size; long seq; int id; } szptr;
static int szptr_cmp(const void *a, const void *b) {
    const szptr *A = (const szptr *)a, *B = (const szptr *)b;
    if (A->size != B->size) return A->size < B->size ? -1 : 1;
    return A->seq < B->seq ? -1 : (A->seq > B->seq ? 1 : 0);
}

/* returns number of result keypoints written to out (coordinates still relative to minBorder) */
static int distribute_octree(const orbo_keypoint *K, int nK, int minX, int maxX, int minY, int maxY, int N,
                             orbo_keypoint *out) {
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return -1;
    const float hX = (float)(maxX - minX) / nIni;
    olist L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    int *ini = (int *)malloc(sizeof(int) * nIni);
    for (int i = 0; i < nIni; i++) {
        int id = ol_new(&L);
        L.n[id].x0 = (int)(hX * (float)i); L.n[id].x1 = (int)(hX * (float)(i + 1));
        L.n[id].y0 = 0; L.n[id].y1 = maxY - minY;
        L.n[id].keys = (int *)malloc(sizeof(int) * (nK ? nK : 1));
        ol_push_back(&L, id);
        ini[i] = id;
    }
    for (int i = 0; i < nK; i++) {
        int b = (int)(K[i].x / hX);
        if (b >= nIni) b = nIni - 1; /* cannot happen for in-range x; guards the index */
        onode *nd = &L.n[ini[b]];
        nd->keys[nd->nkeys++] = i;
    }
    free(ini);
    for (int it = L.head; it >= 0;) {
        if (L.n[it].nkeys == 1) { L.n[it].no_more = 1; it = L.n[it].next; }
        else if (L.n[it].nkeys == 0) it = ol_erase(&L, it);
        else it = L.n[it].next;
    }
    int finish = 0;
    szptr *vsz = NULL, *vprev = NULL; int nsz = 0, szcap = 0, prevcap = 0;
#define VSZ_PUSH(ID) do { if (nsz == szcap) { szcap = szcap ? szcap * 2 : 1024; vsz = (szptr *)realloc(vsz, sizeof(szptr) * szcap); } \
        vsz[nsz].size = L.n[ID].nkeys; vsz[nsz].seq = L.n[ID].seq; vsz[nsz].id = (ID); nsz++; } while (0)
    while (!finish) {
        int prevSize = L.size;
        int nToExpand = 0;
        nsz = 0;
        for (int it = L.head; it >= 0;) {
            if (L.n[it].no_more) { it = L.n[it].next; continue; }
            int c[4];
            divide_node(&L, it, K, c);
            for (int q = 0; q < 4; q++) {
                if (c[q] < 0) continue;
                ol_push_front(&L, c[q]);
                if (L.n[c[q]].nkeys > 1) { nToExpand++; VSZ_PUSH(c[q]); }
            }
            it = ol_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) {
            finish = 1;
        } else if (L.size + nToExpand * 3 > N) {
            while (!finish) {
                prevSize = L.size;
                if (nsz > prevcap) { prevcap = nsz; vprev = (szptr *)realloc(vprev, sizeof(szptr) * (prevcap ? prevcap : 1)); }
                int nprev = nsz;
                if (nprev) memcpy(vprev, vsz, sizeof(szptr) * nprev);
                nsz = 0;
                qsort(vprev, nprev, sizeof(szptr), szptr_cmp);
                for (int j = nprev - 1; j >= 0; j--) {
                    int c[4];
                    divide_node(&L, vprev[j].id, K, c);
                    for (int q = 0; q < 4; q++) {
                        if (c[q] < 0) continue;
                        ol_push_front(&L, c[q]);
                        if (L.n[c[q]].nkeys > 1) VSZ_PUSH(c[q]);
                    }
                    ol_erase(&L, vprev[j].id);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
#undef VSZ_PUSH
    int nout = 0;
    for (int it = L.head; it >= 0; it = L.n[it].next) {
        const onode *nd = &L.n[it];
        int best = nd->keys[0];
        float maxr = K[best].response;
        for (int k = 1; k < nd->nkeys; k++)
            if (K[nd->keys[k]].response > maxr) { best = nd->keys[k]; maxr = K[best].response; }
        out[nout++] = K[best];
    }
    for (int i = 0; i < L.count; i++) free(L.n[i].keys);
    free(L.n); free(vsz); free(vprev);
    return nout;
}

/* IC_Angle, :77-104; center points at the keypoint pixel of the un-blurred level */
static const int k_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
float orbo_ic_angle(const uint8_t *center, int step) {
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH; u <= HALF_PATCH; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH; ++v) {
        int v_sum = 0, d = k_umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orbo_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor, :108-147 */
void orbo_descriptor(const uint8_t *center, int step, float angle_deg, uint8_t desc[32]) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    float angle = angle_deg * factorPI;
    float a, b;
    orbo_sincos_f(angle, &b, &a);
    const int8_t *p = k_pattern;
    for (int i = 0; i < 32; i++) {
        int val = 0;
        for (int bit = 0; bit < 8; bit++, p += 4) {
            const int x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
            int t0 = center[orbo_cv_round_f(x0 * b + y0 * a) * step + orbo_cv_round_f(x0 * a - y0 * b)];
            int t1 = center[orbo_cv_round_f(x1 * b + y1 * a) * step + orbo_cv_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << bit;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ORBextractor::operator(), :1043-1105 */
int orbo_extract(orbo_extractor *e, const uint8_t *img, int w, int h, int stride, orbo_keypoint *kps,
                 uint8_t *desc, int cap) {
    memset(e->t_stage, 0, sizeof(e->t_stage));
    if (!img || w <= 0 || h <= 0) return 0; /* `if(_image.empty()) return;` */
    double t0 = now_s();
    if (compute_pyramid(e, img, w, h, stride)) return -1;
    e->t_stage[0] = now_s() - t0;
    for (int l = 0; l < e->nlevels; l++) {
        for (int v = 0; v <= HALF_PATCH; v++) if (e->umax[v] != k_umax[v]) return -2;
    }
    int total = 0;
    int level_off[MAX_LEVELS + 1];
    orbo_keypoint *tmp = (orbo_keypoint *)malloc(sizeof(orbo_keypoint) * (size_t)(orbo_extractor_capacity(e) + 8));
    for (int l = 0; l < e->nlevels; l++) {
        t0 = now_s();
        fast_cells(e, l);
        e->t_stage[1] += now_s() - t0;
        t0 = now_s();
        const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
        const int maxBX = e->lw[l] - EDGE_THRESHOLD + 3, maxBY = e->lh[l] - EDGE_THRESHOLD + 3;
        level_off[l] = total;
        if ((int)roundf((float)(maxBX - minBX) / (maxBY - minBY)) > NINI_MAX) { free(tmp); return -3; }
        int n = distribute_octree(e->cand[l], e->ncand[l], minBX, maxBX, minBY, maxBY, e->quota[l], tmp + total);
        if (n < 0) { free(tmp); return -3; }
        const int scaledPatch = (int)(PATCH_SIZE * e->scale[l]);
        for (int i = 0; i < n; i++) {
            orbo_keypoint *k = &tmp[total + i];
            k->x += minBX; k->y += minBY; k->octave = l; k->size = (float)scaledPatch;
        }
        total += n;
        e->t_stage[2] += now_s() - t0;
    }
    level_off[e->nlevels] = total;
    if (total > cap) { free(tmp); return -4; }
    /* orientation on the un-blurred pyramid, :472-479 */
    t0 = now_s();
    for (int l = 0; l < e->nlevels; l++) {
        const uint8_t *in = e->pyr[l] + (size_t)EDGE_THRESHOLD * e->lstride[l] + EDGE_THRESHOLD;
        for (int i = level_off[l]; i < level_off[l + 1]; i++) {
            orbo_keypoint *k = &tmp[i];
            k->angle = orbo_ic_angle(in + (size_t)orbo_cv_round_f(k->y) * e->lstride[l] + orbo_cv_round_f(k->x),
                                     e->lstride[l]);
        }
    }
    e->t_stage[3] = now_s() - t0;
    /* blur + descriptors + rescale, :1075-1104 */
    for (int l = 0; l < e->nlevels; l++) {
        const int n = level_off[l + 1] - level_off[l];
        if (n == 0) continue;
        t0 = now_s();
        const int lw = e->lw[l], lh = e->lh[l];
        uint8_t *work = (uint8_t *)malloc((size_t)lw * lh);
        const uint8_t *in = e->pyr[l] + (size_t)EDGE_THRESHOLD * e->lstride[l] + EDGE_THRESHOLD;
        orbo_gaussian7_u8(in, lw, lh, e->lstride[l], work, lw);
        e->t_stage[4] += now_s() - t0;
        t0 = now_s();
        for (int i = level_off[l]; i < level_off[l + 1]; i++) {
            orbo_keypoint *k = &tmp[i];
            orbo_descriptor(work + (size_t)orbo_cv_round_f(k->y) * lw + orbo_cv_round_f(k->x), lw, k->angle,
                            desc + (size_t)i * 32);
            if (l != 0) { k->x *= e->scale[l]; k->y *= e->scale[l]; }
        }
        e->t_stage[5] += now_s() - t0;
        free(work);
    }
    memcpy(kps, tmp, sizeof(orbo_keypoint) * total);
    free(tmp);
    return total;
}

/* test hook: DistributeOctTree on an arbitrary candidate list (coordinates relative to minX/minY) */
int orbo_distribute(const orbo_keypoint *K, int nK, int minX, int maxX, int minY, int maxY, int N,
                    orbo_keypoint *out) {
    return distribute_octree(K, nK, minX, maxX, minY, maxY, N, out);
}

int orbo_level_info(const orbo_extractor *e, int level, int *w, int *h, int *stride) {
    if (level < 0 || level >= e->nlevels || !e->pyr[level]) return -1;
    *w = e->lw[level]; *h = e->lh[level]; *stride = e->lstride[level];
    return 0;
}
const uint8_t *orbo_level_ptr(const orbo_extractor *e, int level) {
    return e->pyr[level] + (size_t)EDGE_THRESHOLD * e->lstride[level] + EDGE_THRESHOLD;
}
int orbo_level_candidates(const orbo_extractor *e, int level, orbo_keypoint *out, int cap) {
    int n = e->ncand[level];
    if (out) memcpy(out, e->cand[level], sizeof(orbo_keypoint) * (n < cap ? n : cap));
    return n;
}
void orbo_stage_seconds(const orbo_extractor *e, double out[6]) { memcpy(out, e->t_stage, sizeof(e->t_stage)); }
