/* orbx CPU oracle, matcher part — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * Restates, in dependency-free C and in the reference's own sequential order:
 *   ORBmatcher::DescriptorDistance                               src/ORBmatcher.cc:1647-1663
 *   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea  src/Frame.cc:259-274, :411-421, :356-409
 *   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)          src/ORBmatcher.cc:45-129
 *   ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, mono)  src/ORBmatcher.cc:1328-1470
 *   ORBmatcher::ComputeThreeMaxima                                          src/ORBmatcher.cc:1601-1642
 * cv::Mat float products are plain sequential float arithmetic (verified against cv2.gemm):
 * ((r0*x0 + r1*x1) + r2*x2) + t.  Compile with -ffp-contract=off.
 * PARITY PINNING: the reference holds no tests or vectors for this path.  PINNED against the reference's own code run here:
 * src/ORBmatcher.cc, Frame.cc, MapPoint.cc and KeyFrame.cc compile unmodified against the OpenCV stand-in oracle/cvmini into
 * oracle/_ref/liborbmatcher_ref.so (make ref), and tests/test_oracle_ref_matcher.py gets identical results from the reference's
 * DescriptorDistance, AssignFeaturesToGrid + GetFeaturesInArea, SearchByProjection(Cur, Last) (all forward / backward / mono
 * modes), SearchByProjection(Frame, MapPoints), the relocalisation SearchByProjection(Cur, KF, sAlreadyFound), SearchByBoW x2,
 * SearchForTriangulation and SearchForInitialization.  Still unpinned by the reference: the loop-closing / Fuse / Sim3 window
 * searches (orbo_match_window), whose oracle entry point takes projections and gates that the adapter evaluates on the host.  tests/test_match_oracle.py
 * additionally cross-checks every function against an independent numpy brute-force statement of the same rules.
 */
#include "orbx_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GRID_COLS 64
#define GRID_ROWS 48
#define TH_HIGH 100
#define TH_LOW 50
#define HISTO_LENGTH 30

int orbo_hamming256(const uint8_t *a, const uint8_t *b) {
    const uint32_t *pa = (const uint32_t *)a, *pb = (const uint32_t *)b;
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* ---- Frame grid ---- */
typedef struct {
    int *start;   /* GRID_COLS*GRID_ROWS + 1, cell = ix*GRID_ROWS + iy */
    int *idx;     /* keypoint indices, ascending inside a cell (push_back order) */
} grid_t;

static void grid_build(const orbo_frame *F, grid_t *g) {
    const int nc = GRID_COLS * GRID_ROWS;
    g->start = (int *)calloc(nc + 1, sizeof(int));
    g->idx = (int *)malloc(sizeof(int) * (F->n > 0 ? F->n : 1));
    int *cell = (int *)malloc(sizeof(int) * (F->n > 0 ? F->n : 1));
    for (int i = 0; i < F->n; i++) {
        /* PosInGrid, Frame.cc:411-421 */
        const int px = (int)roundf((F->keys_un[i].x - F->min_x) * F->grid_w_inv);
        const int py = (int)roundf((F->keys_un[i].y - F->min_y) * F->grid_h_inv);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) { cell[i] = -1; continue; }
        cell[i] = px * GRID_ROWS + py;
        g->start[cell[i] + 1]++;
    }
    for (int c = 0; c < nc; c++) g->start[c + 1] += g->start[c];
    int *cur = (int *)malloc(sizeof(int) * nc);
    memcpy(cur, g->start, sizeof(int) * nc);
    for (int i = 0; i < F->n; i++) if (cell[i] >= 0) g->idx[cur[cell[i]]++] = i;
    free(cur); free(cell);
}
static void grid_free(grid_t *g) { free(g->start); free(g->idx); }

/* GetFeaturesInArea, Frame.cc:356-409; returns count, indices in out (cap F->n) */
static int features_in_area(const orbo_frame *F, const grid_t *g, float x, float y, float r, int minLevel, int maxLevel,
                            int *out) {
    int n = 0;
    int nMinCellX = (int)floorf((x - F->min_x - r) * F->grid_w_inv); if (nMinCellX < 0) nMinCellX = 0;
    if (nMinCellX >= GRID_COLS) return 0;
    int nMaxCellX = (int)ceilf((x - F->min_x + r) * F->grid_w_inv); if (nMaxCellX > GRID_COLS - 1) nMaxCellX = GRID_COLS - 1;
    if (nMaxCellX < 0) return 0;
    int nMinCellY = (int)floorf((y - F->min_y - r) * F->grid_h_inv); if (nMinCellY < 0) nMinCellY = 0;
    if (nMinCellY >= GRID_ROWS) return 0;
    int nMaxCellY = (int)ceilf((y - F->min_y + r) * F->grid_h_inv); if (nMaxCellY > GRID_ROWS - 1) nMaxCellY = GRID_ROWS - 1;
    if (nMaxCellY < 0) return 0;
    const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int c = ix * GRID_ROWS + iy;
            for (int j = g->start[c]; j < g->start[c + 1]; j++) {
                const orbo_keypoint *kp = &F->keys_un[g->idx[j]];
                if (bCheckLevels) {
                    if (kp->octave < minLevel) continue;
                    if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                }
                const float distx = kp->x - x, disty = kp->y - y;
                if (fabsf(distx) < r && fabsf(disty) < r) out[n++] = g->idx[j];
            }
        }
    return n;
}

int orbo_features_in_area(const orbo_frame *F, float x, float y, float r, int minLevel, int maxLevel, int *out) {
    grid_t g;
    grid_build(F, &g);
    const int n = features_in_area(F, &g, x, y, r, minLevel, maxLevel, out);
    grid_free(&g);
    return n;
}

/* ORBmatcher.cc:1601-1642 on bin sizes */
void orbo_three_maxima(const int *histo, int L, int *ind1, int *ind2, int *ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    *ind1 = *ind2 = *ind3 = -1;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

/* SearchByProjection(Frame &F, const vector<MapPoint*>&, th), ORBmatcher.cc:45-129.
 * match[k] (F->n entries) is the caller's mvpMapPoints as point indices: -1 = none on entry unless F->claimed[k];
 * on return match[k] = index of the map point written there, untouched otherwise. */
int orbo_search_by_projection_points(const orbo_frame *F, int n_pts, const orbo_track_point *P, const uint8_t *pt_desc,
                                     float th, float nnratio, int32_t *match) {
    grid_t g;
    grid_build(F, &g);
    int *ind = (int *)malloc(sizeof(int) * (F->n > 0 ? F->n : 1));
    uint8_t *blocked = (uint8_t *)calloc(F->n > 0 ? F->n : 1, 1);   /* mvpMapPoints[idx] && Observations()>0 */
    for (int k = 0; k < F->n; k++) blocked[k] = F->claimed ? F->claimed[k] : 0;
    int nmatches = 0;
    const int bFactor = th != 1.0f;
    for (int i = 0; i < n_pts; i++) {
        const orbo_track_point *p = &P[i];
        if (!p->in_view) continue;                 /* !mbTrackInView || isBad() */
        const int lvl = p->level;
        float r = p->view_cos > 0.998f ? 2.5f : 4.0f;   /* RadiusByViewingCos, :131-137 */
        if (bFactor) r *= th;
        const float rs = r * F->scale_factors[lvl];
        const int n = features_in_area(F, &g, p->proj_x, p->proj_y, rs, lvl - 1, lvl, ind);
        if (n == 0) continue;
        const uint8_t *d = pt_desc + (size_t)32 * i;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int c = 0; c < n; c++) {
            const int idx = ind[c];
            if (blocked[idx]) continue;
            if (F->u_right && F->u_right[idx] > 0) {
                const float er = fabsf(p->proj_xr - F->u_right[idx]);
                if (er > rs) continue;
            }
            const int dist = orbo_hamming256(d, F->desc + (size_t)32 * idx);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F->keys_un[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F->keys_un[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            match[bestIdx] = i;
            blocked[bestIdx] = p->blocks;   /* later points skip it iff this map point has Observations()>0 */
            nmatches++;
        }
    }
    free(ind); free(blocked);
    grid_free(&g);
    return nmatches;
}

/* SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), ORBmatcher.cc:1328-1470.
 * forward / backward are the reference's bForward / bBackward (:1350-1351), computed by the caller from the
 * two poses.  match as above (indices into the last frame). */
int orbo_search_by_projection_frame(const orbo_frame *Cur, int n_last, const orbo_last_point *Lp, const uint8_t *last_desc,
                                    const float Rcw[9], const float tcw[3], int forward, int backward, float th,
                                    int check_ori, int32_t *match) {
    grid_t g;
    grid_build(Cur, &g);
    int *ind = (int *)malloc(sizeof(int) * (Cur->n > 0 ? Cur->n : 1));
    uint8_t *blocked = (uint8_t *)calloc(Cur->n > 0 ? Cur->n : 1, 1);
    for (int k = 0; k < Cur->n; k++) blocked[k] = Cur->claimed ? Cur->claimed[k] : 0;
    int *hist_kp = (int *)malloc(sizeof(int) * (n_last > 0 ? n_last : 1));   /* rotHist entries in push order */
    int *hist_bin = (int *)malloc(sizeof(int) * (n_last > 0 ? n_last : 1));
    int nh = 0, nmatches = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < n_last; i++) {
        const orbo_last_point *p = &Lp[i];
        if (!p->valid) continue;                   /* pMP && !mvbOutlier[i] */
        const float xc = ((Rcw[0] * p->x + Rcw[1] * p->y) + Rcw[2] * p->z) + tcw[0];
        const float yc = ((Rcw[3] * p->x + Rcw[4] * p->y) + Rcw[5] * p->z) + tcw[1];
        const float zc = ((Rcw[6] * p->x + Rcw[7] * p->y) + Rcw[8] * p->z) + tcw[2];
        const float invzc = (float)(1.0 / (double)zc);
        if (invzc < 0) continue;
        const float u = Cur->fx * xc * invzc + Cur->cx;
        const float v = Cur->fy * yc * invzc + Cur->cy;
        if (u < Cur->min_x || u > Cur->max_x) continue;
        if (v < Cur->min_y || v > Cur->max_y) continue;
        const int oct = p->octave;
        const float radius = th * Cur->scale_factors[oct];
        int n;
        if (forward) n = features_in_area(Cur, &g, u, v, radius, oct, -1, ind);
        else if (backward) n = features_in_area(Cur, &g, u, v, radius, 0, oct, ind);
        else n = features_in_area(Cur, &g, u, v, radius, oct - 1, oct + 1, ind);
        if (n == 0) continue;
        const uint8_t *d = last_desc + (size_t)32 * i;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < n; c++) {
            const int i2 = ind[c];
            if (blocked[i2]) continue;
            if (Cur->u_right && Cur->u_right[i2] > 0) {
                const float ur = u - Cur->bf * invzc;
                const float er = fabsf(ur - Cur->u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = orbo_hamming256(d, Cur->desc + (size_t)32 * i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            match[bestIdx2] = i;
            blocked[bestIdx2] = p->blocks;
            nmatches++;
            if (check_ori) {
                float rot = p->angle - Cur->keys_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                hist_kp[nh] = bestIdx2; hist_bin[nh] = bin; nh++;
            }
        }
    }
    if (check_ori) {
        int cnt[HISTO_LENGTH] = {0}, i1, i2, i3;
        for (int j = 0; j < nh; j++) cnt[hist_bin[j]]++;
        orbo_three_maxima(cnt, HISTO_LENGTH, &i1, &i2, &i3);
        for (int j = 0; j < nh; j++)
            if (hist_bin[j] != i1 && hist_bin[j] != i2 && hist_bin[j] != i3) { match[hist_kp[j]] = -1; nmatches--; }
    }
    free(ind); free(blocked); free(hist_kp); free(hist_bin);
    grid_free(&g);
    return nmatches;
}

/* SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist),
 * ORBmatcher.cc:1472-1599 (relocalisation).  The caller evaluates the per-point host geometry exactly as the reference
 * does (isBad / sAlreadyFound / distance-invariance gate :1516-1521 -> valid; MapPoint::PredictScale :1523 -> octave);
 * every non-NULL mvpMapPoints entry blocks (claimed), and so does every match made here. */
int orbo_search_by_projection_kf(const orbo_frame *Cur, int n_pts, const orbo_last_point *Lp, const uint8_t *pt_desc,
                                 const float Rcw[9], const float tcw[3], float th, int orb_dist, int check_ori, int32_t *match) {
    grid_t g;
    grid_build(Cur, &g);
    int *ind = (int *)malloc(sizeof(int) * (Cur->n > 0 ? Cur->n : 1));
    uint8_t *blocked = (uint8_t *)calloc(Cur->n > 0 ? Cur->n : 1, 1);
    for (int k = 0; k < Cur->n; k++) blocked[k] = Cur->claimed ? Cur->claimed[k] : 0;
    int *hist_kp = (int *)malloc(sizeof(int) * (n_pts > 0 ? n_pts : 1)), *hist_bin = (int *)malloc(sizeof(int) * (n_pts > 0 ? n_pts : 1));
    int nh = 0, nmatches = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < n_pts; i++) {
        const orbo_last_point *p = &Lp[i];
        if (!p->valid) continue;
        const float xc = ((Rcw[0] * p->x + Rcw[1] * p->y) + Rcw[2] * p->z) + tcw[0];
        const float yc = ((Rcw[3] * p->x + Rcw[4] * p->y) + Rcw[5] * p->z) + tcw[1];
        const float zc = ((Rcw[6] * p->x + Rcw[7] * p->y) + Rcw[8] * p->z) + tcw[2];
        const float invzc = (float)(1.0 / (double)zc);          /* no depth-sign test in this overload */
        const float u = Cur->fx * xc * invzc + Cur->cx;
        const float v = Cur->fy * yc * invzc + Cur->cy;
        if (u < Cur->min_x || u > Cur->max_x) continue;
        if (v < Cur->min_y || v > Cur->max_y) continue;
        const int lvl = p->octave;                                /* nPredictedLevel */
        const float radius = th * Cur->scale_factors[lvl];
        const int n = features_in_area(Cur, &g, u, v, radius, lvl - 1, lvl + 1, ind);
        if (n == 0) continue;
        const uint8_t *d = pt_desc + (size_t)32 * i;
        int bestDist = 256, bestIdx2 = -1;
        for (int c = 0; c < n; c++) {
            const int i2 = ind[c];
            if (blocked[i2]) continue;
            const int dist = orbo_hamming256(d, Cur->desc + (size_t)32 * i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            match[bestIdx2] = i;
            blocked[bestIdx2] = 1;
            nmatches++;
            if (check_ori) {
                float rot = p->angle - Cur->keys_un[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)roundf(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                hist_kp[nh] = bestIdx2; hist_bin[nh] = bin; nh++;
            }
        }
    }
    if (check_ori) {
        int cnt[HISTO_LENGTH] = {0}, i1, i2, i3;
        for (int j = 0; j < nh; j++) cnt[hist_bin[j]]++;
        orbo_three_maxima(cnt, HISTO_LENGTH, &i1, &i2, &i3);
        for (int j = 0; j < nh; j++)
            if (hist_bin[j] != i1 && hist_bin[j] != i2 && hist_bin[j] != i3) { match[hist_kp[j]] = -1; nmatches--; }
    }
    free(ind); free(blocked); free(hist_kp); free(hist_bin);
    grid_free(&g);
    return nmatches;
}

/* ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:405-520.
 * prev[i] = vbPrevMatched[i1] (x, y); match12 out (F1 keypoints -> F2 index or -1); returns nmatches.  The final
 * "update prev matched" step (:513-516) is a copy the caller does from match12. */
int orbo_search_for_initialization(const orbo_frame *F1, const orbo_frame *F2, const float *prev_xy, int window_size, float nnratio,
                                   int check_ori, int32_t *match12) {
    grid_t g;
    grid_build(F2, &g);
    int *ind = (int *)malloc(sizeof(int) * (F2->n > 0 ? F2->n : 1));
    int *matched_dist = (int *)malloc(sizeof(int) * (F2->n > 0 ? F2->n : 1)), *match21 = (int *)malloc(sizeof(int) * (F2->n > 0 ? F2->n : 1));
    int *hist_i1 = (int *)malloc(sizeof(int) * (F1->n > 0 ? F1->n : 1)), *hist_bin = (int *)malloc(sizeof(int) * (F1->n > 0 ? F1->n : 1));
    for (int k = 0; k < F2->n; k++) { matched_dist[k] = 2147483647; match21[k] = -1; }
    for (int i = 0; i < F1->n; i++) match12[i] = -1;
    int nh = 0, nmatches = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i1 = 0; i1 < F1->n; i1++) {
        const int level1 = F1->keys_un[i1].octave;
        if (level1 > 0) continue;
        const int n = features_in_area(F2, &g, prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)window_size, level1, level1, ind);
        if (n == 0) continue;
        const uint8_t *d1 = F1->desc + (size_t)32 * i1;
        int bestDist = 2147483647, bestDist2 = 2147483647, bestIdx2 = -1;
        for (int c = 0; c < n; c++) {
            const int i2 = ind[c];
            const int dist = orbo_hamming256(d1, F2->desc + (size_t)32 * i2);
            if (matched_dist[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if ((float)bestDist < (float)bestDist2 * nnratio) {
                if (match21[bestIdx2] >= 0) { match12[match21[bestIdx2]] = -1; nmatches--; }
                match12[i1] = bestIdx2;
                match21[bestIdx2] = i1;
                matched_dist[bestIdx2] = bestDist;
                nmatches++;
                if (check_ori) {
                    float rot = F1->keys_un[i1].angle - F2->keys_un[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    hist_i1[nh] = i1; hist_bin[nh] = bin; nh++;
                }
            }
        }
    }
    if (check_ori) {
        int cnt[HISTO_LENGTH] = {0}, a, b, c;
        for (int j = 0; j < nh; j++) cnt[hist_bin[j]]++;
        orbo_three_maxima(cnt, HISTO_LENGTH, &a, &b, &c);
        for (int j = 0; j < nh; j++)
            if (hist_bin[j] != a && hist_bin[j] != b && hist_bin[j] != c && match12[hist_i1[j]] >= 0) { match12[hist_i1[j]] = -1; nmatches--; }
    }
    free(ind); free(matched_dist); free(match21); free(hist_i1); free(hist_bin);
    grid_free(&g);
    return nmatches;
}

/* Window + Hamming core shared by ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (ORBmatcher.cc:290-403),
 * Fuse(KeyFrame*, vpMapPoints, th) (:825-975), Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint) (:977-1100) and the two
 * passes of SearchBySim3 (:1102-1326).  The caller evaluates each point's host geometry in the reference's own arithmetic
 * (projection, IsInImage, distance-invariance and viewing-angle gates -> valid; PredictScale -> the level window;
 * th * scale[level] -> radius) and applies the outcome (vpMatched / map mutations / mutual check) on the host.
 *   flags & 1  reprojection chi2 gate of Fuse (:903-927): 5.99 (monocular keypoint) / 7.8 (stereo keypoint, with ur)
 *   flags & 2  claims: keypoints with F->claimed, or chosen by an earlier point, are skipped (vpMatched[idx], :375-376)
 * best_idx[i] = chosen keypoint or -1 (none, or best distance > max_dist); best_dist[i] = its distance (256 if none). */
int orbo_match_window(const orbo_frame *F, int n_pts, const orbo_window_point *P, const uint8_t *pt_desc, int flags,
                      const float *inv_sigma2, int max_dist, int32_t *best_idx, int32_t *best_dist) {
    grid_t g;
    grid_build(F, &g);
    int *ind = (int *)malloc(sizeof(int) * (F->n > 0 ? F->n : 1));
    uint8_t *blocked = (uint8_t *)calloc(F->n > 0 ? F->n : 1, 1);
    if (flags & 2) for (int k = 0; k < F->n; k++) blocked[k] = F->claimed ? F->claimed[k] : 0;
    int nacc = 0;
    for (int i = 0; i < n_pts; i++) {
        const orbo_window_point *p = &P[i];
        best_idx[i] = -1; best_dist[i] = 256;
        if (!p->valid) continue;
        const int n = features_in_area(F, &g, p->u, p->v, p->radius, -1, -1, ind);     /* KeyFrame::GetFeaturesInArea has no level filter */
        int bestDist = 256, bestIdx = -1;
        for (int c = 0; c < n; c++) {
            const int idx = ind[c];
            if ((flags & 2) && blocked[idx]) continue;
            const orbo_keypoint *kp = &F->keys_un[idx];
            if (kp->octave < p->min_level || kp->octave > p->max_level) continue;
            if (flags & 1) {
                const float ex = p->u - kp->x, ey = p->v - kp->y;
                if (F->u_right && F->u_right[idx] >= 0) {
                    const float er = p->ur - F->u_right[idx];
                    const float e2 = ex * ex + ey * ey + er * er;
                    if (e2 * inv_sigma2[kp->octave] > 7.8) continue;
                } else {
                    const float e2 = ex * ex + ey * ey;
                    if (e2 * inv_sigma2[kp->octave] > 5.99) continue;
                }
            }
            const int dist = orbo_hamming256(pt_desc + (size_t)32 * i, F->desc + (size_t)32 * idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestIdx >= 0 && bestDist <= max_dist) {
            best_idx[i] = bestIdx; best_dist[i] = bestDist;
            if (flags & 2) blocked[bestIdx] = 1;
            nacc++;
        }
    }
    free(ind); free(blocked);
    grid_free(&g);
    return nacc;
}

/* ---- bucket (vocabulary-node) matchers -------------------------------------------------------------------------
 * ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)        src/ORBmatcher.cc:159-288   mode 0
 * ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...)     src/ORBmatcher.cc:522-655   mode 1
 * ORBmatcher::SearchForTriangulation(...)                src/ORBmatcher.cc:657-823   mode 2  (+ CheckDistEpipolarLine :140-157)
 * A DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) is given as sorted node ids + CSR lists.
 * match_a[i] = index in set B matched to feature i of set A, or -1.  For mode 0 the reference's output is indexed by
 * the frame's features (vpMapPointMatches[idxF] = pMP of idxKF); that array is the inverse of match_a (every B index
 * is claimed at most once) and the caller inverts it. */
static int check_epipolar(const orbo_keypoint *kp1, const orbo_keypoint *kp2, const float *F12, const float *sigma2_b) {
    const float a = kp1->x * F12[0] + kp1->y * F12[3] + F12[6];
    const float b = kp1->x * F12[1] + kp1->y * F12[4] + F12[7];
    const float c = kp1->x * F12[2] + kp1->y * F12[5] + F12[8];
    const float num = a * kp2->x + b * kp2->y + c;
    const float den = a * a + b * b;
    if (den == 0) return 0;
    const float dsqr = num * num / den;
    return dsqr < 3.84 * sigma2_b[kp2->octave];
}

int orbo_match_buckets(const orbo_bucket_job *J, int32_t *match_a) {
    const orbo_bow_set *A = &J->a, *B = &J->b;
    uint8_t *matched_b = (uint8_t *)calloc(B->n > 0 ? B->n : 1, 1);
    int *hist_a = (int *)malloc(sizeof(int) * (A->n > 0 ? A->n : 1)), *hist_bin = (int *)malloc(sizeof(int) * (A->n > 0 ? A->n : 1));
    int nh = 0, nmatches = 0;
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < A->n; i++) match_a[i] = -1;
    int ia = 0, ib = 0;
    while (ia < A->n_nodes && ib < B->n_nodes) {
        if (A->node_id[ia] == B->node_id[ib]) {
            for (int k1 = A->node_start[ia]; k1 < A->node_start[ia + 1]; k1++) {
                const int idx1 = A->node_feat[k1];
                if (!A->valid[idx1]) continue;               /* mode 0/1: !pMP || isBad; mode 2: already has a MapPoint */
                const uint8_t *d1 = A->desc + (size_t)32 * idx1;
                if (J->mode == 2) {
                    const int bStereo1 = A->u_right && A->u_right[idx1] >= 0;
                    if (J->only_stereo && !bStereo1) continue;
                    int bestDist = TH_LOW, bestIdx2 = -1;
                    for (int k2 = B->node_start[ib]; k2 < B->node_start[ib + 1]; k2++) {
                        const int idx2 = B->node_feat[k2];
                        if (matched_b[idx2] || !B->valid[idx2]) continue;     /* vbMatched2 is never set by the reference */
                        const int bStereo2 = B->u_right && B->u_right[idx2] >= 0;
                        if (J->only_stereo && !bStereo2) continue;
                        const int dist = orbo_hamming256(d1, B->desc + (size_t)32 * idx2);
                        if (dist > TH_LOW || dist > bestDist) continue;
                        const orbo_keypoint *kp2 = &B->keys_un[idx2];
                        if (!bStereo1 && !bStereo2) {
                            const float distex = J->ex - kp2->x, distey = J->ey - kp2->y;
                            if (distex * distex + distey * distey < 100 * J->scale_b[kp2->octave]) continue;
                        }
                        if (check_epipolar(&A->keys_un[idx1], kp2, J->F12, J->sigma2_b)) { bestIdx2 = idx2; bestDist = dist; }
                    }
                    if (bestIdx2 >= 0) {
                        match_a[idx1] = bestIdx2;
                        nmatches++;
                        if (J->check_ori) {
                            float rot = A->keys_un[idx1].angle - B->keys_un[bestIdx2].angle;
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            hist_a[nh] = idx1; hist_bin[nh] = bin; nh++;
                        }
                    }
                } else {
                    int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                    for (int k2 = B->node_start[ib]; k2 < B->node_start[ib + 1]; k2++) {
                        const int idx2 = B->node_feat[k2];
                        if (matched_b[idx2]) continue;                          /* vpMapPointMatches[realIdxF] / vbMatched2[idx2] */
                        if (J->mode == 1 && !B->valid[idx2]) continue;           /* !pMP2 || pMP2->isBad() */
                        const int dist = orbo_hamming256(d1, B->desc + (size_t)32 * idx2);
                        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                        else if (dist < bestDist2) bestDist2 = dist;
                    }
                    const int pass = J->mode == 0 ? bestDist1 <= TH_LOW : bestDist1 < TH_LOW;
                    if (pass && (float)bestDist1 < J->nnratio * (float)bestDist2) {
                        match_a[idx1] = bestIdx2;
                        matched_b[bestIdx2] = 1;
                        if (J->check_ori) {
                            float rot = A->keys_un[idx1].angle - B->keys_un[bestIdx2].angle;
                            if (rot < 0.0) rot += 360.0f;
                            int bin = (int)roundf(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            hist_a[nh] = idx1; hist_bin[nh] = bin; nh++;
                        }
                        nmatches++;
                    }
                }
            }
            ia++; ib++;
        } else if (A->node_id[ia] < B->node_id[ib]) {
            while (ia < A->n_nodes && A->node_id[ia] < B->node_id[ib]) ia++;     /* lower_bound */
        } else {
            while (ib < B->n_nodes && B->node_id[ib] < A->node_id[ia]) ib++;
        }
    }
    if (J->check_ori) {
        int cnt[HISTO_LENGTH] = {0}, i1, i2, i3;
        for (int j = 0; j < nh; j++) cnt[hist_bin[j]]++;
        orbo_three_maxima(cnt, HISTO_LENGTH, &i1, &i2, &i3);
        for (int j = 0; j < nh; j++)
            if (hist_bin[j] != i1 && hist_bin[j] != i2 && hist_bin[j] != i3) { match_a[hist_a[j]] = -1; nmatches--; }
    }
    free(matched_b); free(hist_a); free(hist_bin);
    return nmatches;
}
