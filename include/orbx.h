/* orbx — C ABI of the B200-native hot path of Active-ORB-SLAM2 (sm_100a CUDA behind plain C).
 *
 * The reference (XinkeAE/Active-ORB-SLAM2) has no FFI layer: its "operator API" for this path is three
 * C++ class surfaces.  Each entry point below names the reference interface it replaces;
 * active-orb-slam2_b200/adapter/ re-creates those C++ classes on top of this ABI, and INTEGRATION.md
 * shows the binding a maintainer adds.
 *
 * Conventions: POD only; every function returns orbx_status (0 = OK, negative = error) and never throws;
 * handles are opaque, independent, and NOT re-entrant (like ORB_SLAM2::ORBextractor, which mutates
 * mvImagePyramid): use one handle per host thread / camera.  `stream` is a cudaStream_t passed as void*
 * (NULL = the legacy default stream).  "_host" entry points take host pointers and are synchronous;
 * "_device" entry points take device pointers and only enqueue work on `stream`.
 */
#ifndef ORBX_H
#define ORBX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int orbx_status;
enum {
    ORBX_OK = 0,
    ORBX_ERR_INVALID = -1,     /* bad argument */
    ORBX_ERR_CUDA = -2,        /* CUDA runtime error; orbx_last_error() has the text */
    ORBX_ERR_NO_DEVICE = -3,   /* no CUDA device / not an sm_100 part */
    ORBX_ERR_CAPACITY = -4,    /* input larger than the handle was created for */
    ORBX_ERR_UNSUPPORTED = -5, /* shape the reference itself cannot process (e.g. nIni == 0) */
    ORBX_ERR_NOMEM = -6,
    ORBX_ERR_ABORTED = -7      /* stop flag was raised (LocalBA) */
};

const char *orbx_last_error(void); /* thread-local text of the last failure */
int orbx_version(void);

/* ---- cv::KeyPoint, 28 bytes (what ORBextractor::operator() fills, ORBextractor.cc:1043) ------------- */
typedef struct {
    float x, y;     /* pt, level-0 pixel coordinates */
    float size;     /* 31 * scale[octave], truncated to int (ORBextractor.cc:837) */
    float angle;    /* degrees, cv::fastAtan2 of the intensity centroid */
    float response; /* FAST score */
    int32_t octave;
    int32_t class_id; /* -1 */
} orbx_keypoint;

/* =====================================================================================================
 * ORBextractor  (reference include/ORBextractor.h:45-111, src/ORBextractor.cc)
 * ===================================================================================================== */
typedef struct orbx_extractor orbx_extractor;

/* replaces ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (ORBextractor.cc:410).  max_width/max_height/max_batch size the device buffers once. */
orbx_status orbx_extractor_create(orbx_extractor **out, int nfeatures, float scale_factor, int nlevels,
                                  int ini_th_fast, int min_th_fast, int max_width, int max_height,
                                  int max_batch, int device);
void orbx_extractor_destroy(orbx_extractor *e);

/* keypoint capacity of ONE frame (sum over levels of quota + slack; DistributeOctTree may return a few
 * more than the quota, ORBextractor.cc:663).  Output arrays are [batch][capacity]. */
int orbx_extractor_capacity(const orbx_extractor *e);

/* replaces GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares / GetInverseScaleSigmaSquares
 * (ORBextractor.h:62-82); any pointer may be NULL; arrays of nlevels. */
orbx_status orbx_extractor_tables(const orbx_extractor *e, float *scale, float *inv_scale, float *sigma2,
                                  float *inv_sigma2, int32_t *features_per_level);

/* replaces ORBextractor::operator()(image, mask, keypoints, descriptors) (ORBextractor.cc:1043) for
 * `batch` independent 8-bit single-channel frames of identical size.
 *   images[b]   host pointer to frame b, `stride` bytes between rows
 *   kps         host, [batch][capacity]        desc  host, [batch][capacity][32]
 *   counts      host, [batch]  number of keypoints of frame b (level-major order like the reference)
 * An empty image (width or height 0) yields counts 0 and ORBX_OK, like the reference's silent return. */
orbx_status orbx_extractor_run_host(orbx_extractor *e, const uint8_t *const *images, int batch, int width,
                                    int height, int stride, orbx_keypoint *kps, uint8_t *desc,
                                    int32_t *counts);

/* same, device-resident: d_images holds `batch` frames `frame_pitch` bytes apart; outputs are device
 * pointers with the layout above.  Only enqueues on `stream`. */
orbx_status orbx_extractor_run_device(orbx_extractor *e, const uint8_t *d_images, size_t frame_pitch,
                                      int batch, int width, int height, int stride, orbx_keypoint *d_kps,
                                      uint8_t *d_desc, int32_t *d_counts, void *stream);

/* replaces the public member mvImagePyramid (ORBextractor.h:85; read by Frame::ComputeStereoMatches,
 * Frame.cc:502,592,609): geometry of level `level` of the last run, and a device pointer to the first
 * INTERIOR pixel of frame `batch_idx` (19-pixel REFLECT_101 pad around it, like the reference). */
orbx_status orbx_extractor_pyramid(const orbx_extractor *e, int batch_idx, int level, const uint8_t **d_ptr,
                                   int *width, int *height, int *pitch);
/* copies that level (interior only, or with the pad when with_border != 0) to host memory, rows `dst_stride` apart */
orbx_status orbx_extractor_pyramid_host(const orbx_extractor *e, int batch_idx, int level, int with_border,
                                        uint8_t *dst, int dst_stride);

/* stage access for parity tests: FAST candidates of one level before DistributeOctTree, as packed words
 * (x | y<<12 | score<<24, coordinates relative to the 16-pixel border like ORBextractor.cc:822-823);
 * order is unspecified.  Returns the count through *n (may exceed cap; only cap are copied). */
orbx_status orbx_extractor_candidates_host(const orbx_extractor *e, int batch_idx, int level, uint32_t *dst,
                                           int cap, int *n);

/* same kind of stage access: the keypoints DistributeOctTree kept for one level, in the reference's list
 * order (ORBextractor.cc:749-768), packed as above; and the 7x7-blurred copy of a level (ORBextractor.cc:1086) */
orbx_status orbx_extractor_level_keypoints_host(const orbx_extractor *e, int batch_idx, int level, uint32_t *dst,
                                                int cap, int *n);
orbx_status orbx_extractor_blurred_host(const orbx_extractor *e, int batch_idx, int level, uint8_t *dst,
                                        int dst_stride);

/* per-stage device timing for bench.py: after orbx_extractor_profile(e, slots) every run records CUDA events
 * on its launching stream around the five stages (pyramid, FAST, quadtree, blur, describe) into a ring of
 * `slots` runs; orbx_extractor_stage_ms sums the elapsed milliseconds of the recorded runs per stage into
 * ms[5], returns how many runs that was, and restarts the ring.  slots = 0 switches it off. */
orbx_status orbx_extractor_profile(orbx_extractor *e, int slots);
orbx_status orbx_extractor_stage_ms(orbx_extractor *e, int *runs, float *ms);

/* number of kernels launched by the last run (bench.py reports it as gpu_launches) */
int orbx_extractor_last_launches(const orbx_extractor *e);

/* =====================================================================================================
 * ORBmatcher  (reference include/ORBmatcher.h:37-102, src/ORBmatcher.cc; Frame grid src/Frame.cc:259-274,
 * :356-421).  The adapter gathers these POD views from Frame / MapPoint with the reference's own getters and
 * writes `match` back into Frame::mvpMapPoints.
 * ===================================================================================================== */

/* replaces static int ORBmatcher::DescriptorDistance(const cv::Mat&, const cv::Mat&) (ORBmatcher.cc:1647);
 * plain host function (256-bit Hamming distance) */
int orbx_hamming256(const uint8_t a[32], const uint8_t b[32]);

typedef struct {                 /* a Frame as the matchers read it (include/Frame.h) */
    int32_t n;                   /* N */
    const int32_t *n_dev;        /* device entry points only: if non-NULL, N is read from here on the device */
    const orbx_keypoint *keys_un; /* mvKeysUn */
    const uint8_t *desc;         /* mDescriptors, n x 32, 16-byte aligned */
    const float *u_right;        /* mvuRight (NULL = all negative) */
    const uint8_t *claimed;      /* 1 where mvpMapPoints[i] && mvpMapPoints[i]->Observations()>0 (NULL = none) */
    float min_x, min_y, max_x, max_y;      /* mnMinX, mnMinY, mnMaxX, mnMaxY */
    float grid_w_inv, grid_h_inv;          /* mfGridElementWidthInv, mfGridElementHeightInv (64 x 48 grid) */
    float fx, fy, cx, cy, bf, b;
    const float *scale_factors;  /* mvScaleFactors */
    int32_t nlevels;
} orbx_frame_view;

typedef struct {                 /* a map point prepared by Frame::isInFrustum (MapPoint.h mTrack* members) */
    float proj_x, proj_y, proj_xr, view_cos;   /* mTrackProjX, mTrackProjY, mTrackProjXR, mTrackViewCos */
    int32_t level;               /* mnTrackScaleLevel */
    uint8_t in_view;             /* mbTrackInView && !isBad() */
    uint8_t blocks;              /* Observations() > 0 */
    uint8_t pad[2];
} orbx_track_point;

typedef struct {                 /* keypoint i of the last frame together with its map point */
    float x, y, z;               /* pMP->GetWorldPos() */
    float angle;                 /* LastFrame.mvKeysUn[i].angle */
    int32_t octave;              /* LastFrame.mvKeys[i].octave */
    uint8_t valid;               /* pMP && !LastFrame.mvbOutlier[i] */
    uint8_t blocks;              /* pMP->Observations() > 0 */
    uint8_t pad[2];
} orbx_last_point;

typedef struct orbx_matcher orbx_matcher;   /* device scratch + stream; ORBmatcher itself is stateless */
orbx_status orbx_matcher_create(orbx_matcher **out, int max_keypoints, int max_points, int max_jobs, int device);
void orbx_matcher_destroy(orbx_matcher *m);

/* replaces int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)
 * (ORBmatcher.cc:45-129), mfNNratio passed as nnratio.  match[F->n] is Frame::mvpMapPoints as indices into
 * `pts` (in/out: entries that receive no match are left untouched); *nmatches is the return value. */
orbx_status orbx_match_projection_points_host(orbx_matcher *m, const orbx_frame_view *F, int n_pts,
                                              const orbx_track_point *pts, const uint8_t *pt_desc, float th,
                                              float nnratio, int32_t *match, int32_t *nmatches);

/* replaces int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th,
 * const bool bMono) (ORBmatcher.cc:1328-1470).  Rcw/tcw = CurrentFrame.mTcw (row-major 3x3, 3); forward /
 * backward = the reference's bForward / bBackward (:1350-1351); check_ori = mbCheckOrientation. */
orbx_status orbx_match_projection_frame_host(orbx_matcher *m, const orbx_frame_view *cur, int n_last,
                                             const orbx_last_point *pts, const uint8_t *last_desc,
                                             const float Rcw[9], const float tcw[3], int forward, int backward,
                                             float th, int check_ori, int32_t *match, int32_t *nmatches);

/* replaces int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound,
 * const float th, const int ORBdist) (ORBmatcher.cc:1472-1599, relocalisation).  pts[i] describes pKF's map point i:
 * valid = pMP && !isBad() && !sAlreadyFound.count(pMP) && the distance-invariance gate of :1516-1521; octave =
 * pMP->PredictScale(dist3D, &CurrentFrame) (:1523) -- both evaluated by the adapter in the reference's own
 * arithmetic; angle = pKF->mvKeysUn[i].angle.  cur->claimed marks every non-NULL mvpMapPoints entry (this overload
 * does not look at Observations()); there is no depth-sign and no uRight test in this overload. */
orbx_status orbx_match_projection_keyframe_host(orbx_matcher *m, const orbx_frame_view *cur, int n_pts,
                                                const orbx_last_point *pts, const uint8_t *pt_desc, const float Rcw[9],
                                                const float tcw[3], float th, int orb_dist, int check_ori,
                                                int32_t *match, int32_t *nmatches);

/* batched, device-resident form of the call above: every pointer inside a job is a device pointer; d_jobs is a
 * device array of n_jobs (<= max_jobs) independent (current, last) pairs.  Only enqueues on `stream`. */
typedef struct {
    orbx_frame_view cur;
    int32_t n_last;
    const orbx_last_point *pts;
    const uint8_t *last_desc;
    float Rcw[9], tcw[3];
    int32_t forward, backward;
    float th;
    int32_t check_ori;
    int32_t *match;              /* cur.n entries, in/out */
    int32_t *nmatches;           /* 1 entry */
    int32_t max_dist;            /* accept threshold: TH_HIGH = 100 for the LastFrame overload; 0 means 100 */
    int32_t variant;             /* 0 = LastFrame overload; 1 = KeyFrame (relocalisation) overload, see below */
} orbx_frame_match_job;
orbx_status orbx_match_projection_frame_device(orbx_matcher *m, const orbx_frame_match_job *d_jobs, int n_jobs,
                                               void *stream);
/* replaces int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, vector<cv::Point2f> &vbPrevMatched,
 * vector<int> &vnMatches12, int windowSize) (ORBmatcher.cc:405-520; monocular initialisation).  prev_xy = vbPrevMatched as
 * (x, y) pairs, F1->n of them; match12 = vnMatches12.  The reference's last step (vbPrevMatched[i1] = F2 keypoint of the
 * match, :513-516) is a copy the adapter does from match12.  A later, closer match may take a keypoint away from an
 * earlier one (:470-474), so the points are processed strictly in order (one warp, lanes over a window's keypoints). */
orbx_status orbx_match_initialization_host(orbx_matcher *m, const orbx_frame_view *F1, const orbx_frame_view *F2,
                                           const float *prev_xy, int window_size, float nnratio, int check_ori,
                                           int32_t *match12, int32_t *nmatches);

/* ---- window + Hamming core of the keyframe projection searches -------------------------------------------------------------
 * serves ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th)  (ORBmatcher.cc:290-403; flags = 2)
 *        ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th)                                      (:825-975;  flags = 1)
 *        ORBmatcher::Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)                    (:977-1100; flags = 0)
 *        both passes of ORBmatcher::SearchBySim3                                           (:1102-1326; flags = 0)
 * The adapter evaluates the per-point host geometry in the reference's own arithmetic (projection, IsInImage, the
 * distance-invariance and viewing-angle gates -> valid; MapPoint::PredictScale -> level window; th * scale -> radius)
 * and applies the outcome on the host in reference order (vpMatched, Replace / AddObservation, the mutual check).
 * F is the KeyFrame seen through orbx_frame_view (KeyFrame::GetFeaturesInArea, KeyFrame.cc:630-669, is the Frame
 * version without the level filter).  flags & 1: reprojection chi2 gate of Fuse (5.99 / 7.8 with inv_sigma2 =
 * mvInvLevelSigma2); flags & 2: claims -- keypoints marked in F->claimed (vpMatched[idx] != NULL) or chosen by an
 * earlier point are skipped.  best_idx[i] = keypoint or -1 (no candidate, or best distance > max_dist);
 * best_dist[i] = its distance (256 if none). */
typedef struct {
    float u, v, ur, radius;
    int32_t min_level, max_level;   /* nPredictedLevel - 1, nPredictedLevel */
    uint8_t valid, pad[3];
} orbx_window_point;
orbx_status orbx_match_window_host(orbx_matcher *m, const orbx_frame_view *F, int n_pts, const orbx_window_point *pts,
                                   const uint8_t *pt_desc, int flags, const float *inv_sigma2, int max_dist,
                                   int32_t *best_idx, int32_t *best_dist, int32_t *n_accepted);

/* ---- vocabulary-node ("bucket") matchers ---------------------------------------------------------------------
 * replaces ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)        (ORBmatcher.cc:159-288)  mode 0
 *          ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)      (ORBmatcher.cc:522-655)  mode 1
 *          ORBmatcher::SearchForTriangulation(KF1, KF2, F12, pairs, onlyStereo)   (ORBmatcher.cc:657-823)  mode 2
 * A DBoW2::FeatureVector (std::map<NodeId, vector<unsigned>>) is passed as ascending node ids + CSR lists.
 * `valid` carries the per-feature map-point tests of the reference: modes 0/1: pMP && !pMP->isBad() (set b: mode 1
 * only); mode 2: the feature has NO map point yet (both sets).
 * match_a[i] = index in set b matched to feature i of set a, or -1.  Mode 0: the reference's output
 * vpMapPointMatches is indexed by the frame's features; it is the inverse of match_a (a frame feature is claimed at
 * most once).  Mode 2: vMatchedPairs = every (i, match_a[i]) with match_a[i] >= 0, in ascending i. */
typedef struct {
    int32_t n;
    const orbx_keypoint *keys_un; /* mvKeysUn */
    const uint8_t *desc;         /* n x 32 */
    const float *u_right;        /* mvuRight (may be NULL) */
    const uint8_t *valid;
    int32_t n_nodes;
    const uint32_t *node_id;     /* ascending */
    const int32_t *node_start;   /* n_nodes + 1 */
    const int32_t *node_feat;    /* feature indices, node by node */
} orbx_bow_set;
typedef struct {
    orbx_bow_set a, b;
    int32_t mode;
    float nnratio;               /* mfNNratio */
    int32_t check_ori, only_stereo;
    float F12[9], ex, ey;        /* mode 2: F12 row-major; epipole of camera 1 in image 2 (ORBmatcher.cc:663-670) */
    const float *sigma2_b, *scale_b;   /* mode 2: pKF2->mvLevelSigma2, pKF2->mvScaleFactors */
    int32_t nlevels;
} orbx_bucket_job;
orbx_status orbx_match_buckets_host(orbx_matcher *m, const orbx_bucket_job *job, int32_t *match_a, int32_t *nmatches);

int orbx_matcher_last_launches(const orbx_matcher *m);
/* diagnostics: how many sweeps the claim resolution of each job of the last call took (synchronises) */
orbx_status orbx_matcher_last_sweeps(orbx_matcher *m, int32_t *out, int n_jobs);

/* =====================================================================================================
 * Frame::ComputeStereoMatches  (reference include/Frame.h:96, src/Frame.cc:495-669).
 * Consumes both keypoint / descriptor sets and both extractors' pyramids (mvImagePyramid) where they already
 * are in device memory: row-band Hamming search, 11x11 SAD refinement over +-5 px on the keypoint's pyramid
 * level, parabola fit, disparity gates, median-distance cut.  Fills mvuRight / mvDepth (-1 = no association).
 * ===================================================================================================== */
typedef struct orbx_stereo orbx_stereo;
/* max_keypoints per image: bound by the search kernel's shared memory (8 bytes per right keypoint: about 28,000 on a B200);
 * ORBX_ERR_CAPACITY beyond it */
orbx_status orbx_stereo_create(orbx_stereo **out, int max_keypoints, int max_pairs, int device);
void orbx_stereo_destroy(orbx_stereo *h);

/* one camera of a batch of rectified pairs, as the extractor left it on the device: pair p has its keypoints at
 * keys + p*pitch, descriptors at desc + p*pitch*32, keypoint count at counts[p*count_step], and its pyramid in slot
 * first_slot + p*slot_step of `extractor`'s last run.  (One extractor run over frames L0 R0 L1 R1 ...: left =
 * {kps, desc, counts, 2*capacity, 2, e, 0, 2}, right = {kps + capacity, desc + 32*capacity, counts + 1, 2*capacity, 2,
 * e, 1, 2}; two extractors as in the reference: pitch = capacity, count_step = 1, first_slot 0, slot_step 1.) */
typedef struct {
    const orbx_keypoint *keys;       /* mvKeys / mvKeysRight (raw keypoints; stereo input is rectified) */
    const uint8_t *desc;             /* mDescriptors / mDescriptorsRight */
    const int32_t *counts;           /* device */
    int32_t pitch, count_step;
    const orbx_extractor *extractor; /* mpORBextractorLeft / mpORBextractorRight */
    int32_t first_slot, slot_step;
    int32_t max_count;               /* upper bound of the counts, sizes the launch (0 = pitch) */
} orbx_stereo_side;

/* replaces void Frame::ComputeStereoMatches() for n_pairs independent pairs; bf = mbf, b = mb.  Outputs (device):
 * d_u_right / d_depth, pair p at + p*out_pitch (out_pitch >= left->pitch); d_kept[p] = associations that survive the
 * median cut (may be NULL).  Only enqueues on `stream` (which must be ordered after the extractor runs). */
orbx_status orbx_stereo_matches_device(orbx_stereo *h, const orbx_stereo_side *left, const orbx_stereo_side *right,
                                       int n_pairs, float bf, float b, float *d_u_right, float *d_depth, int out_pitch,
                                       int32_t *d_kept, void *stream);
/* one pair, host keypoints / descriptors in, host mvuRight / mvDepth out (n_left each); the pyramids are those of slot
 * left_slot / right_slot of the two extractors' last runs.  Synchronous. */
orbx_status orbx_stereo_matches_host(orbx_stereo *h, const orbx_extractor *left, int left_slot, const orbx_extractor *right,
                                     int right_slot, const orbx_keypoint *keys_l, const uint8_t *desc_l, int n_left,
                                     const orbx_keypoint *keys_r, const uint8_t *desc_r, int n_right, float bf, float b,
                                     float *u_right, float *depth, int32_t *n_kept);
/* the same for the pair that the two extractors' orbx_extractor_run_host calls just returned: keypoints, descriptors and counts
 * are read from the extractors' own device buffers (mvKeys / mDescriptors ARE what operator() returned), so nothing is uploaded.
 * n_left = number of left keypoints (the count the left extractor returned for that slot). */
orbx_status orbx_stereo_matches_extractors_host(orbx_stereo *h, const orbx_extractor *left, int left_slot, const orbx_extractor *right,
                                                int right_slot, int n_left, float bf, float b, float *u_right, float *depth,
                                                int32_t *n_kept);
int orbx_stereo_last_launches(const orbx_stereo *h);

/* =====================================================================================================
 * Batched many-sequence mode (BASELINE.json configs 2 and 5): n independent sequences advance in lockstep, one new
 * frame (or rectified pair) per sequence per step, through the per-frame chain of Tracking::TrackWithMotionModel
 * (Tracking.cc:857-880): ORBextractor::operator(), Frame::ComputeStereoMatches (stereo input), and
 * ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono), where the last frame's map points are its own keypoints
 * unprojected with their depth by Frame::UnprojectStereo (Frame.cc:695-709), as Tracking::UpdateLastFrame creates them
 * (Tracking.cc:941-947; here for every keypoint that has a depth).  The previous frame's keypoints, descriptors and depths stay on
 * the device between steps; a step takes HOST images and poses in and gives HOST results back.  Sequences are independent, so
 * handles on one GPU or on several shard them with no exchange.
 * ===================================================================================================== */
typedef struct {
    int32_t nfeatures; float scale_factor; int32_t nlevels, ini_th, min_th;   /* ORBextractor constructor */
    int32_t width, height, n_sequences;
    int32_t stereo;              /* 1: every step brings a left and a right image per sequence (order L0 R0 L1 R1 ...) */
    float fx, fy, cx, cy, bf;    /* Frame::fx ... mbf; mb = bf / fx; the input is undistorted (mnMinX = 0 ... mnMaxX = width) */
    float th;                    /* SearchByProjection's th */
    int32_t check_ori, mono;     /* mbCheckOrientation, bMono */
    float const_depth;           /* stereo == 0 only: the depth (> 0) every keypoint of the last frame is unprojected with (mvDepth of an
                                    RGB-D style input whose scene is a fronto-parallel plane; bench.py's config C2) */
    int32_t device;
    int32_t pose;                /* 1: Optimizer::PoseOptimization of every new frame from its match array, starting at the given Tcw (Tracking.cc:868-870) */
    int32_t n_sub;               /* 0 or 1: the whole step runs on one stream.  > 1: the sequences are cut into n_sub sub-batches that run as
                                    independent pipelines on streams of their own (a sub-batch's next step starts when its own last one is done);
                                    a step's device results are then complete after orbx_sequences_join / orbx_sequences_step_end */
} orbx_sequences_config;
typedef struct {                 /* host pointers, filled when orbx_sequences_step_end returns; capacity = orbx_sequences_capacity() */
    orbx_keypoint *kps;          /* [n_images][capacity]      mvKeys of every new image (may be NULL) */
    uint8_t *desc;               /* [n_images][capacity][32]  mDescriptors (may be NULL) */
    int32_t *counts;             /* [n_images] */
    int32_t *match;              /* [n_sequences][capacity]   CurrentFrame.mvpMapPoints as indices of last-frame keypoints, -1 = none */
    int32_t *nmatches;           /* [n_sequences]             the return value of SearchByProjection (0 at a sequence's first step) */
    float *u_right, *depth;      /* [n_sequences][capacity]   mvuRight / mvDepth of the left image (stereo; may be NULL) */
    double *pose;                /* [n_sequences][7]          config pose: the optimised SE3Quat (quaternion x,y,z,w, translation); untouched at a
                                                               sequence's first step (may be NULL) */
    int32_t *n_inliers;          /* [n_sequences]             PoseOptimization's return value (may be NULL) */
    uint8_t *outlier;            /* [n_sequences][capacity]   mvbOutlier (may be NULL) */
} orbx_sequences_outputs;
typedef struct orbx_sequences orbx_sequences;
orbx_status orbx_sequences_create(orbx_sequences **out, const orbx_sequences_config *cfg);
void orbx_sequences_destroy(orbx_sequences *h);
int orbx_sequences_capacity(const orbx_sequences *h);
/* forget the last frames: the next step is every sequence's first */
orbx_status orbx_sequences_reset(orbx_sequences *h);
/* one step.  images: n_images frames `image_pitch` bytes apart, rows `stride` bytes apart (pinned memory makes the upload
 * asynchronous); Tcw: [n_sequences][12], rows of [Rcw | tcw] of the new frames (the motion-model prediction that
 * SearchByProjection projects with).  _begin enqueues upload, kernels and downloads on the handle's stream and returns; _end waits
 * for them and reports device-side overflow.  One step in flight per handle.  The output arrays may be page-locked (the results
 * are written there by DMA, whole rows) or ordinary memory (the results go to a pinned block of the handle and _end copies the valid
 * part into the arrays: keypoints / descriptors up to counts[i], the per-sequence rows whole); either way they are complete only
 * when _end has returned. */
orbx_status orbx_sequences_step_begin(orbx_sequences *h, const uint8_t *images, size_t image_pitch, int stride, const float *Tcw,
                                      const orbx_sequences_outputs *out);
orbx_status orbx_sequences_step_end(orbx_sequences *h);
orbx_status orbx_sequences_step_host(orbx_sequences *h, const uint8_t *images, size_t image_pitch, int stride, const float *Tcw,
                                     const orbx_sequences_outputs *out);   /* _begin + _end */
/* the same step with the new images already in device memory and the results left there (bench.py's device-resident figure; chaining
 * into orbx_pose_from_matches_device): only enqueues on `stream` (taken as it is: NULL is the legacy default stream;
 * orbx_sequences_device_view gives the handle's own).  The job / pose staging of the handle is
 * rewritten by the next step, so the caller orders steps on one stream.  orbx_sequences_device_view gives the device buffers of the
 * last step (valid until the step after the next one starts) and the extractor of the
 * first sub-batch (stage timing when the handle was created with n_sub = 1; pyramids of its images). */
orbx_status orbx_sequences_step_device(orbx_sequences *h, const uint8_t *d_images, size_t frame_pitch, int stride, const float *Tcw,
                                       void *stream);
typedef struct {
    orbx_extractor *extractor;
    const orbx_keypoint *kps; const uint8_t *desc; const int32_t *counts;      /* [n_images][capacity] ... */
    const int32_t *match, *nmatches; const float *u_right, *depth;             /* [n_sequences][capacity] ... */
    const orbx_frame_match_job *jobs;                                          /* [n_sequences], the jobs of the last projection search */
    const double *pose; const int32_t *n_inliers; const uint8_t *outlier;      /* config pose */
    void *stream;
} orbx_sequences_device;
orbx_status orbx_sequences_device_view(const orbx_sequences *h, orbx_sequences_device *view);
/* replaces the poses the handle keeps of the frames of the last step (it keeps the Tcw that step was given): a caller that optimises
 * poses (config pose, or its own PoseOptimization) hands the optimised Tcw back before the next step, so that the last frame's points
 * are unprojected with them, like Tracking's mLastFrame.  [n_sequences][12] */
orbx_status orbx_sequences_set_last_poses(orbx_sequences *h, const float *Tcw_last);
/* n_sub > 1: makes `stream` wait for every sub-batch's last step (a no-op for n_sub <= 1, where steps run on the caller's stream) */
orbx_status orbx_sequences_join(orbx_sequences *h, void *stream);
int orbx_sequences_last_launches(const orbx_sequences *h);

/* =====================================================================================================
 * Optimizer::LocalBundleAdjustment  (reference include/Optimizer.h:45, src/Optimizer.cc:454-779, and the g2o
 * pieces it drives: types_six_dof_expmap.{h,cpp}, base_binary_edge.hpp:55-120, robust_kernel_impl.cpp:78-91,
 * block_solver.hpp:354-486, optimization_algorithm_levenberg.cpp:61-189).
 * The adapter builds this POD problem exactly where the reference builds the g2o graph (Optimizer.cc:486-655)
 * and writes the result back where it reads the graph (Optimizer.cc:709-778).
 * ===================================================================================================== */
typedef struct {
    int32_t n_kf;                /* keyframe vertices: local (free) and fixed */
    const double *kf_pose;       /* n_kf x 7: quaternion (x,y,z,w) then translation = Converter::toSE3Quat(pKF->GetPose()) */
    const uint8_t *kf_fixed;     /* vSE3->setFixed(...) (Optimizer.cc:529, :541) */
    int32_t n_pts;
    const double *pts;           /* n_pts x 3, Converter::toVector3d(pMP->GetWorldPos()) */
    int32_t n_edges;
    const int32_t *e_kf, *e_pt;  /* vertex indices of every observation */
    const double *e_obs;         /* n_edges x 3: kpUn.pt.x, kpUn.pt.y, mvuRight (third unused for monocular edges) */
    const float *e_inv_sigma2;   /* mvInvLevelSigma2[kpUn.octave] */
    const uint8_t *e_stereo;     /* 1 = EdgeStereoSE3ProjectXYZ, 0 = EdgeSE3ProjectXYZ */
    double fx, fy, cx, cy, bf;
    const volatile uint8_t *stop_flag;   /* pbStopFlag (may be NULL); polled between LM trials */
} orbx_lba_problem;

typedef struct {
    double *kf_pose;             /* n_kf x 7, optimised SE3Quat of every keyframe vertex */
    double *pts;                 /* n_pts x 3 */
    double *chi2;                /* per edge: e->chi2() as the reference reads it at Optimizer.cc:718/:732 (may be NULL) */
    uint8_t *erase;              /* per edge: 1 = goes to vToErase (chi2 over 5.991 / 7.815 or depth not positive) (may be NULL) */
    int32_t lm_trials;           /* Levenberg trials over both rounds */
    int32_t stopped;             /* 1 = the stop flag was already set, nothing was optimised (Optimizer.cc:656-658) */
    /* optional inspection of the very first trial (parity tests; NULL to skip): reduced camera system
     * Hschur (dim x dim, row-major, symmetric), bschur, pose update x_p, with dim = 6 x free keyframes */
    double *first_Hschur, *first_bschur, *first_xp;
    double first_lambda;
} orbx_lba_result;

typedef struct orbx_lba orbx_lba;
orbx_status orbx_lba_create(orbx_lba **out, int max_keyframes, int max_points, int max_edges, int device);
void orbx_lba_destroy(orbx_lba *h);

/* replaces Optimizer::LocalBundleAdjustment from optimizer.initializeOptimization() (Optimizer.cc:659) to the
 * erase list (:735): optimize(its1) with Huber kernels, outlier classification, optimize(its2) without them.
 * The reference calls it with its1 = 5, its2 = 10.  Host pointers; synchronous. */
orbx_status orbx_lba_solve_host(orbx_lba *h, const orbx_lba_problem *prob, int its1, int its2, orbx_lba_result *res);

/* asynchronous form of orbx_lba_solve_host for many independent windows in flight (batched many-sequence mode, SURVEY
 * §8e): one handle per window, each with its own stream.  _begin uploads the window and enqueues everything (both
 * rounds, the outlier passes, the read-back into pinned staging) without waiting; _end waits for that handle and fills
 * `res` (first_* inspection fields are not filled).  The stop flag is looked at once, in _begin.  Only windows that fit
 * the single-kernel path (<= 36 free keyframes, <= 64 keyframes) are accepted; others return ORBX_ERR_UNSUPPORTED. */
orbx_status orbx_lba_solve_begin(orbx_lba *h, const orbx_lba_problem *prob, int its1, int its2);
orbx_status orbx_lba_solve_end(orbx_lba *h, const orbx_lba_problem *prob, orbx_lba_result *res);

/* one Levenberg trial's linear-system work on the current estimates, for bench.py ("LocalBA Schur build"):
 * residuals + Jacobians + quadratic form (BlockSolver::buildSystem) + Schur complement (BlockSolver::solve up to
 * the linear solve) with the given lambda, `reps` times back to back on the handle's stream; the elapsed device
 * time of the repetitions comes back in *ms (CUDA events).  The problem of the last orbx_lba_solve_host call, or
 * `prob` if non-NULL, is used. */
orbx_status orbx_lba_build_schur_timed(orbx_lba *h, const orbx_lba_problem *prob, double lambda, int reps, float *ms,
                                       double *Hschur, double *bschur);
int orbx_lba_last_launches(const orbx_lba *h);
/* diagnostics: nanoseconds the cluster kernel of the last orbx_lba_solve_host spent per phase
 * (quadratic form, Schur accumulation, cluster reduction, reduced solve, update, residuals) */
orbx_status orbx_lba_phase_ns(const orbx_lba *h, double out[6]);

/* =====================================================================================================
 * Optimizer::PoseOptimization  (reference include/Optimizer.h:49, src/Optimizer.cc:239-452): the
 * motion-only bundle adjustment Tracking runs right after every matcher call (Tracking.cc:870, :994, :1039).
 * The adapter lists the keypoints that have a map point in index order, exactly where the reference builds the g2o
 * graph (:275-350), and writes pose / mvbOutlier back where it reads them (:371-417, :424-430).
 * ===================================================================================================== */
typedef struct {
    int32_t n;                   /* nInitialCorrespondences */
    const double *Xw;            /* n x 3: pMP->GetWorldPos() widened to double (:305-308) */
    const double *obs;           /* n x 3: kpUn.pt.x, kpUn.pt.y, mvuRight[i]; a negative third value = monocular edge (:281) */
    const float *inv_sigma2;     /* mvInvLevelSigma2[kpUn.octave] */
    double pose[7];              /* Converter::toSE3Quat(pFrame->mTcw): quaternion (x,y,z,w), translation */
    double fx, fy, cx, cy, bf;   /* the frame's float intrinsics widened to double (e->fx = pFrame->fx ...) */
} orbx_pose_problem;
typedef struct {
    double pose[7];              /* SE3quat_recov (:424-427); Converter::toCvMat + SetPose stay in the adapter */
    uint8_t *outlier;            /* caller-allocated, n entries: mvbOutlier of the listed keypoints */
    int32_t n_inliers;           /* the return value: nInitialCorrespondences - nBad (0 if n < 3, :355) */
    int32_t n_bad;               /* pFrame->nBadPoseOpt */
    int32_t lm_trials;           /* Levenberg trials over the four rounds (diagnostics) */
} orbx_pose_result;

typedef struct orbx_pose orbx_pose;
/* max_observations: over the whole batch of one call */
orbx_status orbx_pose_create(orbx_pose **out, int max_observations, int max_frames, int device);
void orbx_pose_destroy(orbx_pose *h);
/* replaces int Optimizer::PoseOptimization(Frame *pFrame) for n_frames independent frames in one launch (one thread
 * block per frame; the four rounds, their Levenberg loops and the classifications run on the device).  Host pointers;
 * synchronous. */
orbx_status orbx_pose_optimize_host(orbx_pose *h, const orbx_pose_problem *probs, int n_frames, orbx_pose_result *res);
/* device-resident: PoseOptimization of every frame of a batch right after orbx_match_projection_frame_device, from the SAME
 * device job array (the frame's keypoints and mvuRight, the last frame's points, the pose Rcw / tcw that the search used, and the
 * match array it filled).  The observations are listed on the device in ascending keypoint index, like the reference adds its
 * edges; d_inv_sigma2 = mvInvLevelSigma2 (nlevels floats).  Outputs (device): d_pose_out n_frames x 7, d_n_inliers n_frames
 * (may be NULL), d_outlier_kp n_frames x kp_pitch = mvbOutlier per keypoint (0 for keypoints without a map point).  The handle
 * must have been created with max_observations >= n_frames * kp_pitch.  Only enqueues on `stream` (3 launches). */
orbx_status orbx_pose_from_matches_device(orbx_pose *h, const orbx_frame_match_job *d_jobs, int n_frames, const float *d_inv_sigma2,
                                          int nlevels, double fx, double fy, double cx, double cy, double bf, double *d_pose_out,
                                          int32_t *d_n_inliers, uint8_t *d_outlier_kp, int kp_pitch, void *stream);
int orbx_pose_last_launches(const orbx_pose *h);

/* =====================================================================================================
 * DBoW2 TemplatedVocabulary::transform  (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1138-1262, called by
 * Frame::ComputeBoW Frame.cc:286-293 and KeyFrame::ComputeBoW KeyFrame.cc:76-85 with levelsup = 4).
 * The device does the per-feature descent of the vocabulary tree (FORB::distance = 256-bit Hamming, the FIRST child
 * among equal distances wins) and returns, per feature, the word id, the word's weight and the node passed at level
 * L - levelsup.  The adapter then runs the reference's own bookkeeping loop over these (BowVector::addWeight /
 * FeatureVector::addFeature in feature order, skipping weight 0, then the L1 normalisation), which produces the
 * std::map objects SearchByBoW / SearchForTriangulation consume.
 * ===================================================================================================== */
typedef struct orbx_vocabulary orbx_vocabulary;
/* the tree as TemplatedVocabulary holds it in m_nodes (node 0 = root): CSR of every node's children in the reference's
 * order, descriptors (32 bytes per node), weights, word ids (leaves).  Copied to the device once. */
orbx_status orbx_vocabulary_create(orbx_vocabulary **out, int n_nodes, const int32_t *child_start, const int32_t *children,
                                   const uint8_t *node_desc, const double *node_weight, const int32_t *node_word_id, int L,
                                   int max_features, int device);
void orbx_vocabulary_destroy(orbx_vocabulary *h);
/* one feature set, host pointers; word / node / weight have n entries.  Synchronous. */
orbx_status orbx_vocabulary_transform_host(orbx_vocabulary *h, const uint8_t *desc, int n, int levelsup, int32_t *word,
                                           int32_t *node, double *weight);
/* batched, device-resident (descriptors where the extractor left them): frame f has its descriptors at d_desc + f*pitch*32
 * and counts[f*count_step] of them (<= max_count); outputs at + f*pitch.  Only enqueues on `stream`. */
orbx_status orbx_vocabulary_transform_device(orbx_vocabulary *h, int levelsup, const uint8_t *d_desc, const int32_t *d_counts,
                                             int count_step, int pitch, int max_count, int batch, int32_t *d_word, int32_t *d_node,
                                             double *d_weight, void *stream);
int orbx_vocabulary_last_launches(const orbx_vocabulary *h);

/* =====================================================================================================
 * MapPoint::ComputeDistinctiveDescriptors  (reference include/MapPoint.h:73, src/MapPoint.cc:275-340),
 * for many map points in one launch (LocalMapping calls it per new / fused point, LocalMapping.cc:195, :441, :651-667).
 * The adapter gathers, per map point, the descriptors of its non-bad observations in the order the reference pushes them
 * into vDescriptors (:293-299) and afterwards sets mDescriptor = vDescriptors[best_idx].clone() under mMutexFeatures.
 * ===================================================================================================== */
typedef struct orbx_mappoints orbx_mappoints;
orbx_status orbx_mappoints_create(orbx_mappoints **out, int max_points, int max_descriptors, int device);
void orbx_mappoints_destroy(orbx_mappoints *h);
/* point p owns descriptors start[p] .. start[p+1] of desc (32 bytes each); best_idx[p] = the descriptor with the least
 * median Hamming distance to the point's other descriptors (first among equals), -1 for an empty set; best_median may be NULL */
orbx_status orbx_mappoints_distinctive_host(orbx_mappoints *h, int n_points, const int32_t *start, const uint8_t *desc,
                                            int32_t *best_idx, int32_t *best_median);
int orbx_mappoints_last_launches(const orbx_mappoints *h);

/* =====================================================================================================
 * Frame::isInFrustum  (reference include/Frame.h:80, src/Frame.cc:298-354) for all points of the local
 * map at once (Tracking::SearchLocalPoints, Tracking.cc:1085-1103).  Produces the orbx_track_point records that
 * orbx_match_projection_points_* consume (mbTrackInView, mTrackProjX/Y/XR, mnTrackScaleLevel, mTrackViewCos).
 * pad[0] of an output record is 1 when the predicted level could not be decided safely on the device (logf within a last bit
 * of a level boundary): the adapter then sets level = pMP->PredictScale(cv::norm(P - mOw), this) for that point.
 * ===================================================================================================== */
typedef struct {
    float x, y, z;               /* pMP->GetWorldPos() */
    float nx, ny, nz;            /* pMP->GetNormal() */
    float min_distance, max_distance;   /* mfMinDistance, mfMaxDistance (the 0.8 / 1.2 factors of the getters are applied inside) */
    uint8_t skip;                /* pMP->isBad() || pMP->mnLastFrameSeen == mCurrentFrame.mnId: not evaluated, in_view = 0 */
    uint8_t blocks;              /* Observations() > 0, copied to the output record */
    uint8_t pad[2];
} orbx_frustum_point;
typedef struct {
    float Rcw[9], tcw[3], Ow[3]; /* mRcw (row-major), mtcw, mOw */
    float fx, fy, cx, cy, bf;    /* bf = mbf */
    float min_x, max_x, min_y, max_y;   /* mnMinX, mnMaxX, mnMinY, mnMaxY */
    float log_scale_factor;      /* mfLogScaleFactor */
    int32_t n_levels;            /* mnScaleLevels */
    float viewing_cos_limit;     /* 0.5 at the reference's call site */
} orbx_frustum_frame;
/* host pointers, synchronous, no handle (stream-ordered temporary buffers); *n_ambiguous = number of records with pad[0] = 1 */
orbx_status orbx_frustum_host(const orbx_frustum_frame *F, int n, const orbx_frustum_point *pts, orbx_track_point *out,
                              int32_t *n_ambiguous, int device);
/* device pointers; only enqueues on `stream` (d_n_ambiguous: one int32 on the device) */
orbx_status orbx_frustum_device(const orbx_frustum_frame *F, int n, const orbx_frustum_point *d_pts, orbx_track_point *d_out,
                                int32_t *d_n_ambiguous, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H */
