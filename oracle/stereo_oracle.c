/* orbx CPU oracle, stereo association — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * Restates Frame::ComputeStereoMatches (reference src/Frame.cc:495-669) in the reference's own order:
 *   row table of the right keypoints                      :504-522
 *   per left keypoint: Hamming search along its row       :535-580
 *   11x11 SAD refinement over +-5 px on the pyramid level :582-625
 *   parabola fit, disparity gates, depth                  :630-652
 *   median-distance cut                                   :656-668
 * OpenCV arithmetic on the path: Mat::convertTo(CV_32F), IL - IL(w,w)*ones, cv::norm(IL, IR, NORM_L1).  All operands are
 * integer-valued floats below 2^24, so the L1 norm is an exact integer whatever the accumulation order or width
 * (tests/test_stereo_oracle.py checks the SAD against cv2.norm on float patches).  Compile with -ffp-contract=off.
 * PARITY PINNING: the reference holds no test for this function.  PINNED against the reference's own code run here: its two
 * ORBextractors + Frame::ComputeStereoMatches, compiled unmodified into oracle/_ref/liborbmatcher_ref.so, give bit-identical
 * mvuRight / mvDepth on VGA-, EuRoC- and KITTI-shaped pairs (tests/test_oracle_ref_matcher.py).  Also cross-checked against an
 * independent numpy / cv2 statement of the same rules.
 *
 * Where the reference has undefined behaviour the oracle (and the CUDA path) define it:
 *   - a row-table index outside [0, nRows) (:518-521) is skipped (cannot happen for extractor keypoints: they keep 19 px
 *     x scale from the border and r = 2 x scale);
 *   - a SAD window outside the level (cv::Mat::rowRange / colRange would throw) skips the keypoint;
 *   - an empty vDistIdx (:657 reads element 0 of an empty vector) leaves everything unmatched.
 */
#include "orbx_oracle.h"
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TH_HIGH 100
#define TH_LOW 50

typedef struct { int dist, idx; } dist_idx;
static int cmp_dist_idx(const void *a, const void *b) {
    const dist_idx *p = (const dist_idx *)a, *q = (const dist_idx *)b;   /* std::pair<int,int> ordering */
    if (p->dist != q->dist) return p->dist < q->dist ? -1 : 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

int orbo_stereo_matches(const orbo_stereo_job *J, float *u_right, float *depth, int32_t *best_right, int32_t *sad) {
    const int N = J->n_left, Nr = J->n_right;
    for (int i = 0; i < N; i++) { u_right[i] = -1.0f; depth[i] = -1.0f; }            /* :497-498 */
    if (best_right) for (int i = 0; i < N; i++) best_right[i] = -1;
    if (sad) for (int i = 0; i < N; i++) sad[i] = -1;
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                                      /* :500 */
    const int nRows = J->lvl_h[0];                                                     /* :502 */

    /* row table, :505-522 (CSR instead of vector<vector<size_t>>; ascending iR inside a row) */
    int *row_start = (int *)calloc((size_t)nRows + 1, sizeof(int));
    for (int pass = 0; pass < 2; pass++) {
        int *cur = NULL, *rows = NULL;
        if (pass == 1) {
            for (int r = 0; r < nRows; r++) row_start[r + 1] += row_start[r];
            cur = (int *)malloc(sizeof(int) * ((size_t)nRows + 1));
            memcpy(cur, row_start, sizeof(int) * ((size_t)nRows + 1));
            rows = (int *)malloc(sizeof(int) * (size_t)(row_start[nRows] > 0 ? row_start[nRows] : 1));
        }
        for (int iR = 0; iR < Nr; iR++) {
            const float kpY = J->keys_r[iR].y;
            const float r = 2.0f * J->scale[J->keys_r[iR].octave];
            const int maxr = (int)ceilf(kpY + r);
            const int minr = (int)floorf(kpY - r);
            for (int yi = minr; yi <= maxr; yi++) {
                if (yi < 0 || yi >= nRows) continue;
                if (pass == 0) row_start[yi + 1]++;
                else rows[cur[yi]++] = iR;
            }
        }
        if (pass == 1) {
            free(cur);
            /* the search itself */
            const float minZ = J->b, minD = 0.0f, maxD = J->bf / minZ;               /* :525-527 */
            dist_idx *vDistIdx = (dist_idx *)malloc(sizeof(dist_idx) * (size_t)(N > 0 ? N : 1));
            int nDist = 0;
            for (int iL = 0; iL < N; iL++) {
                const orbo_keypoint *kpL = &J->keys_l[iL];
                const int levelL = kpL->octave;
                const float vL = kpL->y, uL = kpL->x;
                const long rowL = (long)vL;                                            /* vRowIndices[vL]: float -> size_t */
                if (rowL < 0 || rowL >= nRows) continue;
                const int c0 = row_start[rowL], c1 = row_start[rowL + 1];
                if (c0 == c1) continue;                                                /* :542-543 */
                const float minU = uL - maxD, maxU = uL - minD;
                if (maxU < 0) continue;
                int bestDist = TH_HIGH, bestIdxR = 0;
                const uint8_t *dL = J->desc_l + (size_t)iL * 32;
                for (int c = c0; c < c1; c++) {                                        /* :556-578 */
                    const int iR = rows[c];
                    const orbo_keypoint *kpR = &J->keys_r[iR];
                    if (kpR->octave < levelL - 1 || kpR->octave > levelL + 1) continue;
                    const float uR = kpR->x;
                    if (uR >= minU && uR <= maxU) {
                        const int dist = orbo_hamming256(dL, J->desc_r + (size_t)iR * 32);
                        if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
                    }
                }
                if (!(bestDist < thOrbDist)) continue;                                 /* :581 */
                if (best_right) best_right[iL] = bestIdxR;
                const float uR0 = J->keys_r[bestIdxR].x;
                const float scaleFactor = J->inv_scale[levelL];
                const float scaleduL = roundf(kpL->x * scaleFactor);
                const float scaledvL = roundf(kpL->y * scaleFactor);
                const float scaleduR0 = roundf(uR0 * scaleFactor);
                const int w = 5, L = 5;
                const int lw = J->lvl_w[levelL], lh = J->lvl_h[levelL];
                const int yl = (int)scaledvL, xl = (int)scaleduL, xr0 = (int)scaleduR0;
                if (yl - w < 0 || yl + w + 1 > lh || xl - w < 0 || xl + w + 1 > lw) continue;   /* Mat::rowRange would throw */
                const uint8_t *PL = J->lvl_l[levelL]; const int sl = J->pitch_l[levelL];
                const uint8_t *PR = J->lvl_r[levelL]; const int sr = J->pitch_r[levelL];
                const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;     /* :604-607 */
                if (iniu < 0 || endu >= (float)lw) continue;
                if (xr0 - L - w < 0) continue;                                         /* Mat::colRange would throw */
                int bestSad = INT_MAX, bestincR = 0;
                float vDists[11];
                const float cL = (float)PL[(size_t)yl * sl + xl];
                for (int incR = -L; incR <= L; incR++) {                               /* :609-623 */
                    const float cR = (float)PR[(size_t)yl * sr + xr0 + incR];
                    double acc = 0.0;                                                  /* cv::norm accumulates CV_32F in double */
                    for (int dy = -w; dy <= w; dy++)
                        for (int dx = -w; dx <= w; dx++) {
                            const float a = (float)PL[(size_t)(yl + dy) * sl + xl + dx] - cL;
                            const float b = (float)PR[(size_t)(yl + dy) * sr + xr0 + incR + dx] - cR;
                            acc += fabs((double)(a - b));
                        }
                    const float dist = (float)acc;
                    if (dist < (float)bestSad) { bestSad = (int)dist; bestincR = incR; }
                    vDists[L + incR] = dist;
                }
                if (bestincR == -L || bestincR == L) continue;                         /* :625-626 */
                const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
                const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));   /* :633 */
                if (deltaR < -1 || deltaR > 1) continue;
                float bestuR = J->scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);  /* :639 */
                float disparity = uL - bestuR;
                if (disparity >= minD && disparity < maxD) {                           /* :643-652 */
                    if (disparity <= 0) {
                        disparity = 0.01;
                        bestuR = uL - 0.01;                                            /* double arithmetic, narrowed on store */
                    }
                    depth[iL] = J->bf / disparity;
                    u_right[iL] = bestuR;
                    vDistIdx[nDist].dist = bestSad; vDistIdx[nDist].idx = iL; nDist++;
                    if (sad) sad[iL] = bestSad;
                }
            }
            int kept = 0;
            if (nDist > 0) {                                                           /* :656-668 */
                qsort(vDistIdx, (size_t)nDist, sizeof(dist_idx), cmp_dist_idx);
                const float median = (float)vDistIdx[nDist / 2].dist;
                const float thDist = 1.5f * 1.4f * median;
                kept = nDist;
                for (int i = nDist - 1; i >= 0; i--) {
                    if ((float)vDistIdx[i].dist < thDist) break;
                    u_right[vDistIdx[i].idx] = -1; depth[vDistIdx[i].idx] = -1;
                    kept--;
                }
            }
            free(vDistIdx); free(rows); free(row_start);
            return kept;
        }
    }
    return 0;
}
