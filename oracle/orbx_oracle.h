/* orbx CPU oracle — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the reference's hot path (XinkeAE/Active-ORB-SLAM2:
 * src/ORBextractor.cc, src/ORBmatcher.cc, src/Optimizer.cc:454-779 and the g2o / OpenCV
 * arithmetic they call).  Nothing in the product library (active-orb-slam2_b200/csrc,
 * include/orbx.h) links, imports or executes this code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, as the checker / CPU baseline.
 *
 * PARITY PINNING: the reference has no tests, golden vectors or fixtures, and its own build (cmake + OpenCV +
 * Eigen + Pangolin) does not run in this image.
 *  - extractor: PINNED against the reference's own src/ORBextractor.cc, compiled unmodified against the OpenCV
 *    stand-in oracle/cvmini into oracle/_ref/liborbextractor_ref.so (make ref): identical keypoints, descriptors
 *    and pyramid levels (tests/test_oracle_ref_extractor.py, tests/golden/ref_extract_digests.txt).  The OpenCV
 *    primitives behind the stand-in (resize, copyMakeBorder, FAST, GaussianBlur, fastAtan2, and cv2.ORB.compute's
 *    descriptor at octave 0) are checked bit-for-bit against cv2 4.13 (tests/test_oracle_cv2.py).
 *  - matchers, Frame grid, stereo association, isInFrustum, distinctive descriptors, vocabulary transform: PINNED against
 *    the reference's own src/ORBmatcher.cc, Frame.cc, MapPoint.cc, KeyFrame.cc and DBoW2 sources, compiled unmodified into
 *    oracle/_ref/liborbmatcher_ref.so (tests/test_oracle_ref_matcher.py, tests/golden/ref_match.npz); vocabulary bookkeeping
 *    also against oracle/_ref/libdbow2_ref.so.  See each file's header for the exact functions.
 *  - the two optimisers (lba_oracle.c, pose_oracle.c): PINNED against the reference's own src/Optimizer.cc, src/Converter.cc and the
 *    whole vendored Thirdparty/g2o, compiled unmodified against the Eigen stand-in oracle/eigenmini into
 *    oracle/_ref/liboptimizer_ref.so (tests/test_oracle_ref_optimizer.py: leaf arithmetic bit-equal, whole functions equal on
 *    reference-built KeyFrame / MapPoint / Frame objects).  What the stand-in decides and Eigen might decide differently is
 *    rounding only (evaluation order of small products, LDLT ordering), three orders of magnitude below the 1e-4 tolerance.
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

/* ---- matchers (src/ORBmatcher.cc, src/Frame.cc grid) ---- */
typedef struct {                 /* a Frame as the matchers read it (include/Frame.h) */
    int32_t n;                   /* N */
    const orbo_keypoint *keys_un; /* mvKeysUn */
    const uint8_t *desc;         /* mDescriptors, n x 32 */
    const float *u_right;        /* mvuRight (NULL = all negative) */
    const uint8_t *claimed;      /* 1 where mvpMapPoints[i] && mvpMapPoints[i]->Observations()>0 (NULL = none) */
    float min_x, min_y, max_x, max_y;      /* mnMinX, mnMinY, mnMaxX, mnMaxY */
    float grid_w_inv, grid_h_inv;          /* mfGridElementWidthInv, mfGridElementHeightInv */
    float fx, fy, cx, cy, bf, b;
    const float *scale_factors;  /* mvScaleFactors */
    int32_t nlevels;
} orbo_frame;
typedef struct {                 /* a map point prepared by Frame::isInFrustum (MapPoint.h mTrack* members) */
    float proj_x, proj_y, proj_xr, view_cos;
    int32_t level;               /* mnTrackScaleLevel */
    uint8_t in_view;             /* mbTrackInView && !isBad() */
    uint8_t blocks;              /* Observations() > 0 */
    uint8_t pad[2];
} orbo_track_point;
typedef struct {                 /* keypoint i of the last frame with its map point */
    float x, y, z;               /* pMP->GetWorldPos() */
    float angle;                 /* LastFrame.mvKeysUn[i].angle */
    int32_t octave;              /* LastFrame.mvKeys[i].octave */
    uint8_t valid;               /* pMP && !mvbOutlier[i] */
    uint8_t blocks;              /* pMP->Observations() > 0 */
    uint8_t pad[2];
} orbo_last_point;
int orbo_hamming256(const uint8_t *a, const uint8_t *b);
int orbo_features_in_area(const orbo_frame *F, float x, float y, float r, int minLevel, int maxLevel, int *out);
void orbo_three_maxima(const int *histo, int L, int *ind1, int *ind2, int *ind3);
int orbo_search_by_projection_points(const orbo_frame *F, int n_pts, const orbo_track_point *P, const uint8_t *pt_desc,
                                     float th, float nnratio, int32_t *match);
int orbo_search_by_projection_frame(const orbo_frame *Cur, int n_last, const orbo_last_point *Lp, const uint8_t *last_desc,
                                    const float Rcw[9], const float tcw[3], int forward, int backward, float th,
                                    int check_ori, int32_t *match);

int orbo_search_by_projection_kf(const orbo_frame *Cur, int n_pts, const orbo_last_point *Lp, const uint8_t *pt_desc,
                                 const float Rcw[9], const float tcw[3], float th, int orb_dist, int check_ori, int32_t *match);
int orbo_search_for_initialization(const orbo_frame *F1, const orbo_frame *F2, const float *prev_xy, int window_size, float nnratio,
                                   int check_ori, int32_t *match12);
typedef struct {                 /* a point already projected into the keyframe by the caller */
    float u, v, ur, radius;      /* projection, u - bf/z, th * mvScaleFactors[nPredictedLevel] */
    int32_t min_level, max_level; /* nPredictedLevel - 1, nPredictedLevel */
    uint8_t valid, pad[3];
} orbo_window_point;
int orbo_match_window(const orbo_frame *F, int n_pts, const orbo_window_point *P, const uint8_t *pt_desc, int flags,
                      const float *inv_sigma2, int max_dist, int32_t *best_idx, int32_t *best_dist);
typedef struct {                 /* one KeyFrame / Frame as the vocabulary-node matchers read it */
    int32_t n;
    const orbo_keypoint *keys_un; /* mvKeysUn (angle, pt, octave) */
    const uint8_t *desc;         /* n x 32 */
    const float *u_right;        /* mvuRight (may be NULL) */
    const uint8_t *valid;        /* see orbo_match_buckets */
    int32_t n_nodes;             /* DBoW2::FeatureVector: node ids ascending, CSR lists of feature indices */
    const uint32_t *node_id;
    const int32_t *node_start, *node_feat;
} orbo_bow_set;
typedef struct {
    orbo_bow_set a, b;
    int32_t mode;                /* 0 SearchByBoW(KF, F), 1 SearchByBoW(KF, KF), 2 SearchForTriangulation */
    float nnratio;
    int32_t check_ori, only_stereo;
    float F12[9], ex, ey;        /* mode 2: fundamental matrix (row-major) and epipole of camera 1 in image 2 */
    const float *sigma2_b, *scale_b;   /* pKF2->mvLevelSigma2, mvScaleFactors */
} orbo_bucket_job;
int orbo_match_buckets(const orbo_bucket_job *J, int32_t *match_a);

/* ---- stereo association (src/Frame.cc:495-669, Frame::ComputeStereoMatches) ---- */
typedef struct {
    int32_t n_left, n_right;             /* N, mvKeysRight.size() */
    const orbo_keypoint *keys_l, *keys_r; /* mvKeys, mvKeysRight (raw keypoints; stereo input is pre-rectified) */
    const uint8_t *desc_l, *desc_r;      /* mDescriptors, mDescriptorsRight */
    int32_t nlevels;
    const float *scale, *inv_scale;      /* mvScaleFactors, mvInvScaleFactors */
    const uint8_t *const *lvl_l;         /* mpORBextractorLeft->mvImagePyramid[l]: first interior pixel of level l */
    const uint8_t *const *lvl_r;         /* mpORBextractorRight->mvImagePyramid[l] */
    const int32_t *lvl_w, *lvl_h, *pitch_l, *pitch_r;
    float bf, b;                         /* mbf, mb */
} orbo_stereo_job;
/* fills mvuRight / mvDepth (n_left each); optional best_right[i] = right keypoint the Hamming search chose (-1 none) and
 * sad[i] = the SAD pushed into vDistIdx (-1 if none).  Returns the number of associations kept after the median cut. */
int orbo_stereo_matches(const orbo_stereo_job *J, float *u_right, float *depth, int32_t *best_right, int32_t *sad);

/* ---- local bundle adjustment (src/Optimizer.cc:454-779 + g2o) ---- */
typedef struct {
    int32_t n_kf;                /* keyframe vertices: local (free) and fixed */
    const double *kf_pose;       /* n_kf x 7: quaternion (x,y,z,w) then translation, = SE3Quat of Converter::toSE3Quat(Tcw) */
    const uint8_t *kf_fixed;     /* vSE3->setFixed(...) */
    int32_t n_pts;
    const double *pts;           /* n_pts x 3, Converter::toVector3d(GetWorldPos()) */
    int32_t n_edges;
    const int32_t *e_kf, *e_pt;  /* vertex indices of every observation */
    const double *e_obs;         /* n_edges x 3: kpUn.pt.x, kpUn.pt.y, mvuRight (third unused for monocular edges) */
    const float *e_inv_sigma2;   /* mvInvLevelSigma2[kpUn.octave] */
    const uint8_t *e_stereo;     /* 1 = EdgeStereoSE3ProjectXYZ, 0 = EdgeSE3ProjectXYZ */
    double fx, fy, cx, cy, bf;
    const volatile uint8_t *stop_flag;   /* pbStopFlag (may be NULL) */
} orbo_lba_problem;
#define ORBO_LBA_MAX_TRACE 256
typedef struct {
    int32_t n_trials;            /* LM trials over both rounds */
    double chi2[ORBO_LBA_MAX_TRACE], lambda[ORBO_LBA_MAX_TRACE];
    int32_t dim;                 /* 6 x free poses */
    double lambda0;
    double *Hschur, *bschur, *xp;   /* optional: reduced system and pose update of the very first trial (dim x dim, dim, dim) */
    double t_build, t_schur, t_solve; /* seconds: errors + quadratic form (per iteration); Schur complement; dense solve + back-substitution (per trial) */
    int32_t n_builds;
} orbo_lba_trace;
/* kf_out n_kf x 7, pt_out n_pts x 3, chi2_out / erase_out per edge (Optimizer.cc:709-735); returns 1 if stopped before starting */
int orbo_lba_solve(const orbo_lba_problem *P, int its1, int its2, double *kf_out, double *pt_out, double *chi2_out,
                   uint8_t *erase_out, orbo_lba_trace *tr);

/* ---- bag-of-words transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1216-1262) ---- */
typedef struct {
    int32_t n_nodes;             /* m_nodes.size(); node 0 is the root */
    const int32_t *child_start;  /* n_nodes + 1: CSR of m_nodes[id].children, in the reference's child order */
    const int32_t *children;
    const uint8_t *desc;         /* n_nodes x 32: m_nodes[id].descriptor (root unused) */
    const double *weight;        /* m_nodes[id].weight */
    const int32_t *word_id;      /* m_nodes[id].word_id (leaves) */
    int32_t L;                   /* m_L */
} orbo_vocabulary;
/* per feature: word id, weight of the leaf reached and the node passed at level L - levelsup (0 if that level is <= 0) */
void orbo_bow_transform(const orbo_vocabulary *V, const uint8_t *desc, int n, int levelsup, int32_t *word, int32_t *node,
                        double *weight);

/* ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:275-340), many map points at once ---- */
/* point p owns descriptors start[p] .. start[p+1] of desc (32 bytes each, in the order the reference pushes them into
 * vDescriptors); best_idx[p] = index inside the point's own set (-1 for an empty set); best_median optional */
void orbo_distinctive_descriptors(int n_points, const int32_t *start, const uint8_t *desc, int32_t *best_idx, int32_t *best_median);

/* ---- Frame::isInFrustum (src/Frame.cc:298-354) ---- */
typedef struct {
    float x, y, z, nx, ny, nz, min_distance, max_distance;   /* GetWorldPos, GetNormal, mfMinDistance, mfMaxDistance */
    uint8_t skip, blocks, pad[2];
} orbo_frustum_point;
typedef struct {
    float Rcw[9], tcw[3], Ow[3], fx, fy, cx, cy, bf, min_x, max_x, min_y, max_y, log_scale_factor;
    int32_t n_levels;
    float viewing_cos_limit;
} orbo_frustum_frame;
void orbo_is_in_frustum(const orbo_frustum_frame *F, int n, const orbo_frustum_point *pts, orbo_track_point *out);
int orbo_predict_scale(float max_distance, float dist, float log_scale_factor, int n_levels);

/* ---- pose-only optimisation (src/Optimizer.cc:239-452 + g2o unary edges) ---- */
typedef struct {
    int32_t n;                   /* keypoints with a map point (nInitialCorrespondences) */
    const double *Xw;            /* n x 3: pMP->GetWorldPos() (float values widened, :305-308) */
    const double *obs;           /* n x 3: kpUn.pt.x, kpUn.pt.y, mvuRight (negative = monocular edge) */
    const float *inv_sigma2;     /* mvInvLevelSigma2[kpUn.octave] */
    double pose[7];              /* Converter::toSE3Quat(pFrame->mTcw): quaternion (x,y,z,w), translation */
    double fx, fy, cx, cy, bf;
} orbo_pose_problem;
/* returns nInitialCorrespondences - nBad; outlier[n] = mvbOutlier of the observations; pose_out = the recovered SE3Quat */
int orbo_pose_optimize(const orbo_pose_problem *P, double pose_out[7], uint8_t *outlier, int32_t *n_bad, int32_t *lm_trials);

/* ---- leaf evaluators of the two optimiser oracles (lba_oracle.c / pose_oracle.c): the same static functions the solvers use, on
 * one edge / vertex; tests/test_oracle_ref_optimizer.py compares them with the reference's g2o classes (oracle/_ref/liboptimizer_ref.so).
 * Jacobians are row-major: A = dE/dX (D x 3), B / J = dE/dxi (D x 6), D = 2 (monocular) or 3 (stereo). */
void orbo_lba_edge_eval(int stereo, const double pose[7], const double X[3], const double obs[3], const double K[5], float inv_sigma2,
                        double *err, double *chi2, int *depth_positive, double *A, double *B);
void orbo_pose_edge_eval(const double pose[7], const double X[3], const double obs[3], const double K[5], float inv_sigma2,
                         double *err, double *chi2, int *depth_positive, double *J);
void orbo_se3_oplus(const double pose[7], const double update[6], double out[7]);   /* VertexSE3Expmap::oplusImpl */
void orbo_se3_map(const double pose[7], const double X[3], double out[3]);          /* SE3Quat::map */
void orbo_huber(double e2, double delta, double rho[3]);                            /* RobustKernelHuber::robustify */
void orbo_to_se3quat(const float Tcw[16], double pose[7]);                          /* Converter::toSE3Quat(cv::Mat) */
void orbo_to_cvmat(const double pose[7], float Tcw[16]);                            /* Converter::toCvMat(SE3Quat) */

#ifdef __cplusplus
}
#endif
#endif
