/* orbx CPU oracle, Frame::isInFrustum — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * Restates reference src/Frame.cc:298-354 with MapPoint::GetMin/MaxDistanceInvariance (MapPoint.cc:427-441) and
 * MapPoint::PredictScale (MapPoint.cc:444-459), one map point at a time, in the reference's float arithmetic:
 *   mRcw*P+mtcw       cv::Mat CV_32F product = ((r0*x0 + r1*x1) + r2*x2) + t, no FMA (checked against cv2.gemm)
 *   cv::norm(PO)      sqrt of the double-accumulated sum of squares, narrowed to float (§8c: checked against cv2.norm)
 *   PO.dot(Pn)        cv::Mat::dot on three floats: products and sum in double (OpenCV's dotProd_32f has no SIMD body for
 *                     fewer than 4 elements) -- NOT reachable from Python, so this one is restated from the OpenCV source
 *   PredictScale      ceil(log(ratio)/mfLogScaleFactor) with float arguments = logf / ceilf (MapPoint.cc includes <math.h>)
 * Compile with -ffp-contract=off.  logf is libm's.
 * PARITY PINNING: the reference holds no test for this function.  PINNED against the reference's own Frame::isInFrustum +
 * MapPoint::PredictScale run from source (oracle/_ref/liborbmatcher_ref.so): bit-equal records on 16 k map points
 * (tests/test_oracle_ref_matcher.py::test_is_in_frustum_equals_the_reference).
 */
#include "orbx_oracle.h"
#include <math.h>

void orbo_is_in_frustum(const orbo_frustum_frame *F, int n, const orbo_frustum_point *pts, orbo_track_point *out) {
    for (int i = 0; i < n; i++) {
        const orbo_frustum_point *p = &pts[i];
        orbo_track_point t;
        t.proj_x = t.proj_y = t.proj_xr = t.view_cos = 0.f;
        t.level = 0; t.in_view = 0; t.blocks = p->blocks; t.pad[0] = t.pad[1] = 0;
        out[i] = t;
        if (p->skip) continue;
        const float *R = F->Rcw;
        const float PcX = ((R[0] * p->x + R[1] * p->y) + R[2] * p->z) + F->tcw[0];
        const float PcY = ((R[3] * p->x + R[4] * p->y) + R[5] * p->z) + F->tcw[1];
        const float PcZ = ((R[6] * p->x + R[7] * p->y) + R[8] * p->z) + F->tcw[2];
        if (PcZ < 0.0f) continue;
        const float invz = 1.0f / PcZ;
        const float u = F->fx * PcX * invz + F->cx;
        const float v = F->fy * PcY * invz + F->cy;
        if (u < F->min_x || u > F->max_x) continue;
        if (v < F->min_y || v > F->max_y) continue;
        const float maxDistance = 1.2f * p->max_distance, minDistance = 0.8f * p->min_distance;
        const float ox = p->x - F->Ow[0], oy = p->y - F->Ow[1], oz = p->z - F->Ow[2];
        const float dist = (float)sqrt((double)ox * ox + (double)oy * oy + (double)oz * oz);
        if (dist < minDistance || dist > maxDistance) continue;
        const float viewCos = (float)(((double)ox * p->nx + (double)oy * p->ny + (double)oz * p->nz) / dist);
        if (viewCos < F->viewing_cos_limit) continue;
        const float ratio = p->max_distance / dist;
        int nScale = (int)ceilf(logf(ratio) / F->log_scale_factor);
        if (nScale < 0) nScale = 0;
        else if (nScale >= F->n_levels) nScale = F->n_levels - 1;
        t.in_view = 1;
        t.proj_x = u;
        t.proj_xr = u - F->bf * invz;
        t.proj_y = v;
        t.level = nScale;
        t.view_cos = viewCos;
        out[i] = t;
    }
}

/* MapPoint::PredictScale alone (for the points the device flags as undecided) */
int orbo_predict_scale(float max_distance, float dist, float log_scale_factor, int n_levels) {
    const float ratio = max_distance / dist;
    int nScale = (int)ceilf(logf(ratio) / log_scale_factor);
    if (nScale < 0) nScale = 0;
    else if (nScale >= n_levels) nScale = n_levels - 1;
    return nScale;
}
