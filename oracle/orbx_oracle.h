/* orbx CPU oracle — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's hot path (XinkeAE/Active-ORB-SLAM2:
 * src/ORBextractor.cc, src/ORBmatcher.cc, src/Optimizer.cc:454-779 and the g2o / OpenCV
 * arithmetic they call).  Nothing in the product library (active-orb-slam2_b200/csrc,
 * include/orbx.h) links, imports or executes this code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, as the checker / CPU baseline.
 *
 * PARITY PINNING: the reference has no tests, golden vectors or fixtures and cannot be compiled
 * in this image (needs OpenCV + Eigen + Pangolin headers) => "parity unpinned" by the reference
 * itself.  What IS pinned (tests/test_oracle_cv2.py): every OpenCV primitive the extractor
 * delegates to (resize, copyMakeBorder, FAST, GaussianBlur, fastAtan2, ORB descriptor of
 * cv2.ORB.compute at octave 0) is checked bit-for-bit against cv2 4.13.
 */
#ifndef ORBX_ORACLE_H
#define ORBX_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* same 28-byte layout as cv::KeyPoint */
typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orbo_keypoint;

typedef struct orbo_extractor orbo_extractor;

/* ---- extractor (src/ORBextractor.cc) ---- */
orbo_extractor *orbo_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th);
void orbo_extractor_destroy(orbo_extractor *e);
void orbo_extractor_tables(const orbo_extractor *e, float *scale, float *inv_scale, float *sigma2,
                           float *inv_sigma2, int *quota, int *umax16);
int orbo_extractor_capacity(const orbo_extractor *e);
/* full operator(): returns number of keypoints (<= cap) or <0 on error */
int orbo_extract(orbo_extractor *e, const uint8_t *img, int w, int h, int stride, orbo_keypoint *kps,
                 uint8_t *desc, int cap);
/* stage access after orbo_extract (for stage-level parity tests) */
int orbo_level_info(const orbo_extractor *e, int level, int *w, int *h, int *stride);
const uint8_t *orbo_level_ptr(const orbo_extractor *e, int level); /* interior origin of the padded buffer */
int orbo_level_candidates(const orbo_extractor *e, int level, orbo_keypoint *out, int cap);
/* DistributeOctTree alone (ORBextractor.cc:539-763); out must hold max(N+3, 4*nIni) entries; <0 = nIni==0 */
int orbo_distribute(const orbo_keypoint *K, int nK, int minX, int maxX, int minY, int maxY, int N,
                    orbo_keypoint *out);
/* seconds spent per stage in the last orbo_extract: pyramid, fast, octree, orient, blur, desc */
void orbo_stage_seconds(const orbo_extractor *e, double out[6]);

/* ---- OpenCV primitives restated (pinned against cv2 in tests) ---- */
void orbo_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh,
                           int dstride);
void orbo_border_reflect101(uint8_t *buf, int w, int h, int stride, int pad); /* buf = interior origin */
void orbo_gaussian7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
int orbo_fast9(const uint8_t *img, int w, int h, int stride, int threshold, int *xs, int *ys, int *scores,
               int cap);
float orbo_fast_atan2(float y, float x);
void orbo_sincos_f(float x, float *s, float *c);
int orbo_cv_round_f(float v);
void orbo_descriptor(const uint8_t *blurred_center, int stride, float angle_deg, uint8_t desc[32]);
float orbo_ic_angle(const uint8_t *center, int stride);

#ifdef __cplusplus
}
#endif
#endif
