// C entry points around the REFERENCE's own ORB_SLAM2::ORBmatcher, Frame and MapPoint (src/ORBmatcher.cc, src/Frame.cc,
// src/MapPoint.cc, src/KeyFrame.cc, src/ORBextractor.cc and DBoW2's BowVector / FeatureVector, compiled unmodified from where they
// lie under /root/reference against the OpenCV stand-in oracle/cvmini; oracle/cvmini/shadow/Converter.h replaces the one header that
// needs Eigen).  TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref_matcher.py drives it to pin oracle/match_oracle.c against the
// reference's code.
//
// The entry points take the oracle's own input records (orbx_oracle.h: orbo_frame, orbo_last_point, orbo_track_point) and
// build the reference's objects from them: a default-constructed Frame whose public members are filled in and whose grid is
// built by the reference's AssignFeaturesToGrid(); MapPoints made by the reference's constructor from a frame row (world
// position, descriptor) with nObs / mTrack* set.  What runs afterwards — the search loops, Frame::GetFeaturesInArea, PosInGrid,
// DescriptorDistance, ComputeThreeMaxima, the pose algebra on cv::Mat — is the reference's code.
#include <opencv2/core/core.hpp>                   // every standard header first: the two defines below must not reach them
#include <mutex>
#include <thread>
#include <cstring>
#include <map>
#include <set>
#include "Thirdparty/DBoW2/DBoW2/BowVector.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"
#include "ORBVocabulary.h"
// the shim fills in state that the reference only reaches through its map / keyframe graph (a map point's normal, distance range
// and descriptor).  Access specifiers do not change layout or mangling, so the reference's own objects are unaffected.
#define protected public
#define private public
#include "ORBmatcher.h"
#undef protected
#undef private
#include "Frame.h"
#include "KeyFrame.h"
#include "KeyFrameDatabase.h"
#include "Map.h"
#include "MapPoint.h"
#include "Converter.h"
#include <cstring>
#include <map>
#include "ref_bump_alloc.h"
#include "orbx_oracle.h"

namespace ORB_SLAM2 {
// Converter.cc:31-39 needs nothing but cv::Mat; the rest of that file needs Eigen
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat &Descriptors) {
    std::vector<cv::Mat> v;
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
}
// src/Map.cc and src/KeyFrameDatabase.cc are not built (Map.cc needs Eigen through Converter.h).  MapPoint::Replace ends in
// mpMap->EraseMapPoint(this) (MapPoint.cc:247), which only removes the point from the map's own set: nothing the tests look at
static void not_built(const char *what) { std::fprintf(stderr, "orbmref: %s is not part of this build\n", what); std::abort(); }
void Map::EraseMapPoint(MapPoint *) {}
void Map::EraseKeyFrame(KeyFrame *) { not_built("Map::EraseKeyFrame"); }
void KeyFrameDatabase::erase(KeyFrame *) { not_built("KeyFrameDatabase::erase"); }
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

namespace {
Map *the_map() {                                   // only its creation mutex is ever touched (MapPoint.cc:55,69,98)
    static Map *m = static_cast<Map *>(std::calloc(1, sizeof(Map)));
    return m;
}
cv::Mat pose4(const float R[9], const float t[3]) {
    cv::Mat T = cv::Mat::eye(4, 4, CV_32F);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T.at<float>(i, j) = R[3 * i + j]; T.at<float>(i, 3) = t[i]; }
    return T;
}
const float kI[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, kZ[3] = {0, 0, 0};

// a Frame as Frame.cc's constructors leave it, from the oracle's record
Frame *make_frame(const orbo_frame *F) {
    Frame *f = new Frame();
    f->N = F->n;
    f->mvKeys.resize(F->n);
    if (F->n) std::memcpy(f->mvKeys.data(), F->keys_un, sizeof(cv::KeyPoint) * F->n);
    f->mvKeysUn = f->mvKeys;
    f->mDescriptors = cv::Mat(F->n, 32, CV_8U, const_cast<uint8_t *>(F->desc)).clone();
    f->mvuRight.assign(F->n, -1.f);
    if (F->u_right) f->mvuRight.assign(F->u_right, F->u_right + F->n);
    f->mvDepth.assign(F->n, -1.f);
    f->mvpMapPoints.assign(F->n, static_cast<MapPoint *>(NULL));
    f->mvbOutlier.assign(F->n, false);
    Frame::fx = F->fx; Frame::fy = F->fy; Frame::cx = F->cx; Frame::cy = F->cy;
    Frame::invfx = 1.0f / F->fx; Frame::invfy = 1.0f / F->fy;
    f->mbf = F->bf; f->mb = F->b;
    Frame::mnMinX = F->min_x; Frame::mnMinY = F->min_y; Frame::mnMaxX = F->max_x; Frame::mnMaxY = F->max_y;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);     // Frame.cc:52-53
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    Frame::mbInitialComputations = false;
    f->mnScaleLevels = F->nlevels;
    f->mvScaleFactors.assign(F->scale_factors, F->scale_factors + F->nlevels);
    f->mfScaleFactor = F->nlevels > 1 ? F->scale_factors[1] : 1.2f;
    f->mfLogScaleFactor = std::log(f->mfScaleFactor);
    f->mvInvScaleFactors.resize(F->nlevels); f->mvLevelSigma2.resize(F->nlevels); f->mvInvLevelSigma2.resize(F->nlevels);
    for (int l = 0; l < F->nlevels; l++) {
        f->mvInvScaleFactors[l] = 1.0f / f->mvScaleFactors[l];
        f->mvLevelSigma2[l] = f->mvScaleFactors[l] * f->mvScaleFactors[l];
        f->mvInvLevelSigma2[l] = 1.0f / f->mvLevelSigma2[l];
    }
    f->mnId = Frame::nNextId++;
    f->SetPose(pose4(kI, kZ));
    f->AssignFeaturesToGrid();                     // the reference's (Frame.cc:259-274)
    return f;
}
// a frame that only carries n descriptor rows (+ octaves) for MapPoint's constructor to copy from (MapPoint.cc:76-100)
Frame *make_carrier(int n, const uint8_t *desc, const orbo_frame *like) {
    Frame *f = new Frame();
    f->N = n;
    f->mvKeysUn.assign(n, cv::KeyPoint());
    f->mDescriptors = cv::Mat(n, 32, CV_8U, const_cast<uint8_t *>(desc)).clone();
    f->mnScaleLevels = like->nlevels;
    f->mvScaleFactors.assign(like->scale_factors, like->scale_factors + like->nlevels);
    f->mnId = Frame::nNextId++;
    f->SetPose(pose4(kI, kZ));
    return f;
}
MapPoint *make_point(float x, float y, float z, Frame *carrier, int row, int n_obs) {
    cv::Mat P = (cv::Mat_<float>(3, 1) << x, y, z);
    MapPoint *p = new MapPoint(P, the_map(), carrier, row);
    p->nObs = n_obs;
    return p;
}
// keypoints of `f` that already hold a map point with observations (the oracle's `claimed`)
void claim(Frame *f, const orbo_frame *F, Frame *carrier, std::vector<MapPoint *> &owned) {
    if (!F->claimed) return;
    for (int k = 0; k < F->n; k++)
        if (F->claimed[k]) { owned.push_back(make_point(1, 1, 1, carrier, 0, 1)); f->mvpMapPoints[k] = owned.back(); }
}
}  // namespace

extern "C" {
int orbmref_hamming256(const uint8_t *a, const uint8_t *b) {
    return ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, const_cast<uint8_t *>(a)), cv::Mat(1, 32, CV_8U, const_cast<uint8_t *>(b)));
}

// Frame::GetFeaturesInArea (Frame.cc:356-409) on the grid built by Frame::AssignFeaturesToGrid
int orbmref_features_in_area(const orbo_frame *F, float x, float y, float r, int min_level, int max_level, int *out) {
    orbref_arena_retain();
    Frame *f = make_frame(F);
    const std::vector<size_t> v = f->GetFeaturesInArea(x, y, r, min_level, max_level);
    for (size_t i = 0; i < v.size(); i++) out[i] = (int)v[i];
    const int n = (int)v.size();
    delete f;
    orbref_arena_release();
    return n;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono), ORBmatcher.cc:1328-1470.
// Tlw (the last frame's pose) only decides bForward / bBackward (:1346-1351); bMono is passed through.
int orbmref_search_by_projection_frame(const orbo_frame *Cur, int n_last, const orbo_last_point *Lp, const uint8_t *last_desc,
                                       const float Rcw[9], const float tcw[3], const float Rlw[9], const float tlw[3], int mono,
                                       float th, float nnratio, int check_ori, int32_t *match) {
    orbref_arena_retain();
    Frame *cur = make_frame(Cur);
    cur->SetPose(pose4(Rcw, tcw));
    // the last frame: one keypoint per record, carrying the descriptor its map point was made from
    orbo_frame L = *Cur;
    std::vector<orbo_keypoint> keys(n_last > 0 ? n_last : 1);
    for (int i = 0; i < n_last; i++) { orbo_keypoint k = {0, 0, 31.f, Lp[i].angle, 0, Lp[i].octave, -1}; keys[i] = k; }
    L.n = n_last; L.keys_un = keys.data(); L.desc = last_desc; L.u_right = NULL; L.claimed = NULL;
    Frame *last = make_frame(&L);
    std::vector<MapPoint *> owned;
    std::map<MapPoint *, int> index_of;
    for (int i = 0; i < n_last; i++) {
        if (!Lp[i].valid) continue;                // pMP == NULL or mvbOutlier[i]: alternate between the two
        MapPoint *p = make_point(Lp[i].x, Lp[i].y, Lp[i].z, last, i, Lp[i].blocks ? 1 : 0);
        owned.push_back(p);
        last->mvpMapPoints[i] = p;
        index_of[p] = i;
    }
    for (int i = 0, flip = 0; i < n_last; i++)
        if (!Lp[i].valid && (flip ^= 1)) {         // every other invalid record: a map point flagged as an outlier (:1358)
            MapPoint *p = make_point(Lp[i].x, Lp[i].y, Lp[i].z, last, i, 1);
            owned.push_back(p);
            last->mvpMapPoints[i] = p;
            last->mvbOutlier[i] = true;
        }
    claim(cur, Cur, last, owned);
    last->SetPose(pose4(Rlw, tlw));
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByProjection(*cur, *last, th, mono != 0);
    for (int k = 0; k < Cur->n; k++) {
        std::map<MapPoint *, int>::const_iterator it = index_of.find(cur->mvpMapPoints[k]);
        match[k] = it == index_of.end() ? -1 : it->second;
    }
    for (size_t i = 0; i < owned.size(); i++) delete owned[i];
    delete last;
    delete cur;
    orbref_arena_release();
    return n;
}

// ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th), ORBmatcher.cc:45-129
int orbmref_search_by_projection_points(const orbo_frame *F, int n_pts, const orbo_track_point *P, const uint8_t *pt_desc, float th,
                                        float nnratio, int32_t *match) {
    orbref_arena_retain();
    Frame *f = make_frame(F);
    Frame *carrier = make_carrier(n_pts > 0 ? n_pts : 1, pt_desc, F);
    std::vector<MapPoint *> pts, owned;
    std::map<MapPoint *, int> index_of;
    for (int i = 0; i < n_pts; i++) {
        MapPoint *p = make_point(0, 0, 1 + i, carrier, i, P[i].blocks ? 1 : 0);
        p->mbTrackInView = P[i].in_view != 0;
        p->mnTrackScaleLevel = P[i].level;
        p->mTrackViewCos = P[i].view_cos;
        p->mTrackProjX = P[i].proj_x; p->mTrackProjY = P[i].proj_y; p->mTrackProjXR = P[i].proj_xr;
        pts.push_back(p);
        index_of[p] = i;
    }
    claim(f, F, carrier, owned);
    ORBmatcher matcher(nnratio, true);
    const int n = matcher.SearchByProjection(*f, pts, th);
    for (int k = 0; k < F->n; k++) {
        std::map<MapPoint *, int>::const_iterator it = index_of.find(f->mvpMapPoints[k]);
        match[k] = it == index_of.end() ? -1 : it->second;
    }
    for (size_t i = 0; i < pts.size(); i++) delete pts[i];
    for (size_t i = 0; i < owned.size(); i++) delete owned[i];
    delete carrier;
    delete f;
    orbref_arena_release();
    return n;
}

// Frame::ComputeStereoMatches (Frame.cc:495-669) after the reference's two ORBextractors have run on the rectified pair, as
// Frame's stereo constructor does (Frame.cc:95-130, without its two threads).  Outputs: the left keypoints / descriptors,
// mvuRight and mvDepth.  Returns the number of left keypoints.
int orbmref_stereo(const uint8_t *left, const uint8_t *right, int w, int h, int stride, int nfeatures, float scale_factor, int nlevels,
                   int ini_th, int min_th, float bf, float b, void *keys_left, uint8_t *desc_left, int cap, float *u_right, float *depth) {
    orbref_arena_retain();
    ORBextractor *el = new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
    ORBextractor *er = new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
    Frame *f = new Frame();
    f->mpORBextractorLeft = el; f->mpORBextractorRight = er;
    (*el)(cv::Mat(h, w, CV_8U, const_cast<uint8_t *>(left), (size_t)stride), cv::Mat(), f->mvKeys, f->mDescriptors);
    (*er)(cv::Mat(h, w, CV_8U, const_cast<uint8_t *>(right), (size_t)stride), cv::Mat(), f->mvKeysRight, f->mDescriptorsRight);
    f->N = (int)f->mvKeys.size();
    f->mvKeysUn = f->mvKeys;
    f->mnScaleLevels = el->GetLevels();
    f->mvScaleFactors = el->GetScaleFactors();
    f->mvInvScaleFactors = el->GetInverseScaleFactors();
    f->mbf = bf; f->mb = b;
    f->ComputeStereoMatches();
    const int n = f->N;
    if (n <= cap) {
        std::memcpy(keys_left, f->mvKeys.data(), sizeof(cv::KeyPoint) * n);
        for (int i = 0; i < n; i++) std::memcpy(desc_left + 32 * (size_t)i, f->mDescriptors.ptr(i), 32);
        std::memcpy(u_right, f->mvuRight.data(), sizeof(float) * n);
        std::memcpy(depth, f->mvDepth.data(), sizeof(float) * n);
    }
    delete f; delete el; delete er;
    orbref_arena_release();
    return n;
}

// Frame::isInFrustum (Frame.cc:298-354; MapPoint::GetMin/MaxDistanceInvariance, PredictScale MapPoint.cc:427-459) for every record.
// Ow_out = the camera centre the reference derived from the pose (Frame::UpdatePoseMatrices).
void orbmref_is_in_frustum(const orbo_frustum_frame *F, int n, const orbo_frustum_point *pts, orbo_track_point *out, float Ow_out[3]) {
    orbref_arena_retain();
    Frame *f = new Frame();
    Frame::fx = F->fx; Frame::fy = F->fy; Frame::cx = F->cx; Frame::cy = F->cy;
    Frame::mnMinX = F->min_x; Frame::mnMaxX = F->max_x; Frame::mnMinY = F->min_y; Frame::mnMaxY = F->max_y;
    f->mbf = F->bf;
    f->mnScaleLevels = F->n_levels;
    f->mfLogScaleFactor = F->log_scale_factor;
    f->mvScaleFactors.assign(F->n_levels, 1.f);
    f->mnId = Frame::nNextId++;
    f->SetPose(pose4(F->Rcw, F->tcw));
    const cv::Mat Ow = f->GetCameraCenter();
    for (int i = 0; i < 3; i++) Ow_out[i] = Ow.at<float>(i);
    Frame *carrier = new Frame();
    const uint8_t zero[32] = {0};
    carrier->N = 1;
    carrier->mvKeysUn.assign(1, cv::KeyPoint());
    carrier->mDescriptors = cv::Mat(1, 32, CV_8U, const_cast<uint8_t *>(zero)).clone();
    carrier->mnScaleLevels = F->n_levels;
    carrier->mvScaleFactors.assign(F->n_levels, 1.f);
    carrier->SetPose(pose4(kI, kZ));
    for (int i = 0; i < n; i++) {
        orbo_track_point t;
        std::memset(&t, 0, sizeof t);
        t.blocks = pts[i].blocks;
        out[i] = t;
        if (pts[i].skip) continue;                 // the caller's gates (Tracking.cc:1085-1093: isBad / already seen in this frame)
        MapPoint *p = make_point(pts[i].x, pts[i].y, pts[i].z, carrier, 0, pts[i].blocks ? 1 : 0);
        p->mNormalVector = (cv::Mat_<float>(3, 1) << pts[i].nx, pts[i].ny, pts[i].nz);
        p->mfMinDistance = pts[i].min_distance;
        p->mfMaxDistance = pts[i].max_distance;
        if (f->isInFrustum(p, F->viewing_cos_limit)) {
            t.in_view = p->mbTrackInView;
            t.proj_x = p->mTrackProjX; t.proj_y = p->mTrackProjY; t.proj_xr = p->mTrackProjXR;
            t.level = p->mnTrackScaleLevel;
            t.view_cos = p->mTrackViewCos;
            out[i] = t;
        }
        delete p;
    }
    delete carrier;
    delete f;
    orbref_arena_release();
}

// ---- the vocabulary-node matchers on KeyFrames built by the reference's KeyFrame(Frame&, Map*, KeyFrameDatabase*) ----
namespace {
const float kK[6] = {517.306408f, 516.469215f, 318.643040f, 255.313989f, 40.0f, 40.0f / 517.306408f};   // TUM1, like orbx.synth
Frame *frame_from_set(const orbo_bow_set *S, int nlevels, const float *scale, const float *sigma2) {
    orbo_frame F;
    std::memset(&F, 0, sizeof F);
    F.n = S->n; F.keys_un = S->keys_un; F.desc = S->desc; F.u_right = S->u_right;
    F.min_x = 0; F.min_y = 0; F.max_x = 640; F.max_y = 480;
    F.fx = kK[0]; F.fy = kK[1]; F.cx = kK[2]; F.cy = kK[3]; F.bf = kK[4]; F.b = kK[5];
    F.scale_factors = scale; F.nlevels = nlevels;
    Frame *f = make_frame(&F);
    f->mvLevelSigma2.assign(sigma2, sigma2 + nlevels);
    for (int j = 0; j < S->n_nodes; j++)
        for (int k = S->node_start[j]; k < S->node_start[j + 1]; k++) f->mFeatVec.addFeature(S->node_id[j], (unsigned)S->node_feat[k]);
    return f;
}
void give_points(Frame *f, const orbo_bow_set *S, bool where_valid, std::vector<MapPoint *> &owned, std::map<MapPoint *, int> &index_of) {
    for (int i = 0; i < S->n; i++)
        if ((S->valid[i] != 0) == where_valid) {
            MapPoint *p = make_point(1, 1, 1 + i, f, i, 1);
            owned.push_back(p);
            f->mvpMapPoints[i] = p;
            index_of[p] = i;
        }
}
}  // namespace

// mode 0 SearchByBoW(KF, F) :159-288, 1 SearchByBoW(KF, KF) :522-655, 2 SearchForTriangulation :657-823.  match_a as the oracle defines
// it.  Mode 2 places camera 1 at the origin and camera 2 at a translation whose epipole is near (J->ex, J->ey); the epipole the
// reference then derives from the two poses (:663-669) is returned in epipole_out for the oracle to use.
int orbmref_match_buckets(const orbo_bucket_job *J, int nlevels, int32_t *match_a, float epipole_out[2]) {
    orbref_arena_retain();
    Frame *fa = frame_from_set(&J->a, nlevels, J->scale_b, J->sigma2_b), *fb = frame_from_set(&J->b, nlevels, J->scale_b, J->sigma2_b);
    std::vector<MapPoint *> owned;
    std::map<MapPoint *, int> in_a, in_b;
    give_points(fa, &J->a, J->mode != 2, owned, in_a);
    if (J->mode != 0) give_points(fb, &J->b, J->mode != 2, owned, in_b);
    if (J->mode == 2) {
        const float t[3] = {(J->ex - kK[2]) / kK[0], (J->ey - kK[3]) / kK[1], 1.0f};
        fb->SetPose(pose4(kI, t));
    }
    KeyFrame *ka = new KeyFrame(*fa, the_map(), NULL), *kb = J->mode != 0 ? new KeyFrame(*fb, the_map(), NULL) : NULL;
    ORBmatcher matcher(J->nnratio, J->check_ori != 0);
    for (int i = 0; i < J->a.n; i++) match_a[i] = -1;
    int n = 0;
    if (J->mode == 0) {
        std::vector<MapPoint *> m;
        n = matcher.SearchByBoW(ka, *fb, m);
        for (int k = 0; k < (int)m.size(); k++)
            if (m[k]) match_a[in_a.at(m[k])] = k;
    } else if (J->mode == 1) {
        std::vector<MapPoint *> m;
        n = matcher.SearchByBoW(ka, kb, m);
        for (int i = 0; i < (int)m.size(); i++)
            if (m[i]) match_a[i] = in_b.at(m[i]);
    } else {
        cv::Mat F12(3, 3, CV_32F);
        for (int i = 0; i < 9; i++) F12.at<float>(i / 3, i % 3) = J->F12[i];
        std::vector<std::pair<size_t, size_t> > pairs;
        n = matcher.SearchForTriangulation(ka, kb, F12, pairs, J->only_stereo != 0);
        for (size_t i = 0; i < pairs.size(); i++) match_a[pairs[i].first] = (int32_t)pairs[i].second;
        const cv::Mat C2 = kb->GetRotation() * ka->GetCameraCenter() + kb->GetTranslation();    // as :663-669
        const float invz = 1.0f / C2.at<float>(2);
        epipole_out[0] = kb->fx * C2.at<float>(0) * invz + kb->cx;
        epipole_out[1] = kb->fy * C2.at<float>(1) * invz + kb->cy;
    }
    delete ka; delete kb;
    for (size_t i = 0; i < owned.size(); i++) delete owned[i];
    delete fa; delete fb;
    orbref_arena_release();
    return n;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize), ORBmatcher.cc:405-520
int orbmref_search_for_initialization(const orbo_frame *F1, const orbo_frame *F2, const float *prev_xy, int window_size, float nnratio,
                                      int check_ori, int32_t *match12) {
    orbref_arena_retain();
    Frame *f1 = make_frame(F1), *f2 = make_frame(F2);
    std::vector<cv::Point2f> prev(F1->n);
    for (int i = 0; i < F1->n; i++) prev[i] = cv::Point2f(prev_xy[2 * i], prev_xy[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchForInitialization(*f1, *f2, prev, m12, window_size);
    for (int i = 0; i < F1->n; i++) match12[i] = m12[i];
    delete f1; delete f2;
    orbref_arena_release();
    return n;
}

// MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:275-340) for n map points; observation j of point i is row start[i] + j of
// desc.  Keyframe j holds, in row i, the descriptor of point i's j-th observation; mObservations is keyed by KeyFrame*, so its
// iteration order is pointer order = creation order under this library's allocator = j.  best_desc[i] = the chosen descriptor.
void orbmref_distinctive_descriptors(int n, const int32_t *start, const uint8_t *desc, uint8_t *best_desc) {
    orbref_arena_retain();
    int max_obs = 0;
    for (int i = 0; i < n; i++) max_obs = std::max(max_obs, start[i + 1] - start[i]);
    const float sf[1] = {1.f};
    std::vector<orbo_keypoint> keys(std::max(n, 1));
    std::memset(keys.data(), 0, sizeof(orbo_keypoint) * keys.size());
    std::vector<Frame *> frames;
    std::vector<KeyFrame *> kfs;
    std::vector<uint8_t> rows((size_t)std::max(n, 1) * 32);
    for (int j = 0; j < max_obs; j++) {
        for (int i = 0; i < n; i++)
            if (j < start[i + 1] - start[i]) std::memcpy(&rows[(size_t)i * 32], desc + (size_t)(start[i] + j) * 32, 32);
        orbo_frame F;
        std::memset(&F, 0, sizeof F);
        F.n = n; F.keys_un = keys.data(); F.desc = rows.data();
        F.max_x = 640; F.max_y = 480; F.fx = F.fy = 500; F.scale_factors = sf; F.nlevels = 1;
        frames.push_back(make_frame(&F));
        kfs.push_back(new KeyFrame(*frames.back(), the_map(), NULL));
    }
    for (int i = 0; i < n; i++) {
        const int cnt = start[i + 1] - start[i];
        std::memset(best_desc + (size_t)i * 32, 0, 32);
        if (cnt == 0) continue;
        MapPoint *p = make_point(1, 1, 1, frames[0], i, 0);
        for (int j = 0; j < cnt; j++) p->mObservations[kfs[j]] = (size_t)i;
        p->ComputeDistinctiveDescriptors();
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(best_desc + (size_t)i * 32, d.ptr(0), 32);
        delete p;
    }
    for (size_t j = 0; j < kfs.size(); j++) { delete kfs[j]; delete frames[j]; }
    orbref_arena_release();
}

// ---- DBoW2 TemplatedVocabulary<FORB::TDescriptor, FORB> (ORBVocabulary): the reference's loader and transform ----
void *orbvref_load_text(const char *path) {
    orbref_arena_retain();
    ORBVocabulary *v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; orbref_arena_release(); return NULL; }
    return v;
}
void orbvref_destroy(void *h) {
    if (!h) return;
    delete static_cast<ORBVocabulary *>(h);
    orbref_arena_release();
}
// transform(features, BowVector, FeatureVector, levelsup) (TemplatedVocabulary.h:1138-1219) as Frame::ComputeBoW calls it, plus the
// per-feature word id (transform(feature), :1061-1074).  BowVector -> (bow_id, bow_val)[*n_bow]; FeatureVector -> node ids
// fv_id[*n_fv], CSR starts fv_start[*n_fv + 1], feature indices fv_feat.
void orbvref_transform(void *h, const uint8_t *desc, int n, int levelsup, int32_t *word, uint32_t *bow_id, double *bow_val, int32_t *n_bow,
                       uint32_t *fv_id, int32_t *fv_start, uint32_t *fv_feat, int32_t *n_fv) {
    const ORBVocabulary &voc = *static_cast<ORBVocabulary *>(h);
    const std::vector<cv::Mat> feats = Converter::toDescriptorVector(cv::Mat(n, 32, CV_8U, const_cast<uint8_t *>(desc)));
    for (int i = 0; i < n; i++) word[i] = (int32_t)voc.transform(feats[i]);
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    voc.transform(feats, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { bow_id[k] = it->first; bow_val[k] = it->second; }
    *n_bow = k;
    k = 0;
    int at = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++k) {
        fv_id[k] = it->first; fv_start[k] = at;
        for (size_t j = 0; j < it->second.size(); j++) fv_feat[at++] = it->second[j];
    }
    fv_start[k] = at;
    *n_fv = k;
}

// ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, th, ORBdist), :1472-1599
// (relocalisation).  Lp[i].valid = the keyframe holds a good map point there that is not in sAlreadyFound (records with valid == 0
// alternate between NULL, bad and already-found).  dist_range[2 i .. 2 i + 1] = mfMinDistance, mfMaxDistance of point i.
// gate_out[i] / level_out[i] = what the adapter evaluates on the host for the oracle's entry point: the distance-invariance gate
// (:1514-1521) and MapPoint::PredictScale (:1523), computed here by the reference's own members.
int orbmref_search_by_projection_kf(const orbo_frame *Cur, int n_pts, const orbo_last_point *Lp, const uint8_t *pt_desc,
                                    const float *dist_range, const float Rcw[9], const float tcw[3], float th, int orb_dist, float nnratio,
                                    int check_ori, int32_t *match, uint8_t *gate_out, int32_t *level_out) {
    orbref_arena_retain();
    Frame *cur = make_frame(Cur);
    cur->SetPose(pose4(Rcw, tcw));
    orbo_frame K = *Cur;
    std::vector<orbo_keypoint> keys(n_pts > 0 ? n_pts : 1);
    for (int i = 0; i < n_pts; i++) { orbo_keypoint k = {0, 0, 31.f, Lp[i].angle, 0, 0, -1}; keys[i] = k; }
    K.n = n_pts; K.keys_un = keys.data(); K.desc = pt_desc; K.u_right = NULL; K.claimed = NULL;
    Frame *kframe = make_frame(&K);
    std::vector<MapPoint *> owned;
    std::map<MapPoint *, int> index_of;
    std::set<MapPoint *> found;
    const cv::Mat R = cur->mTcw.rowRange(0, 3).colRange(0, 3), t = cur->mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -R.t() * t;                                  // as :1478
    for (int i = 0, kind = 0; i < n_pts; i++) {
        gate_out[i] = 0; level_out[i] = 0;
        if (!Lp[i].valid && (kind = (kind + 1) % 3) == 0) continue;  // NULL entry
        MapPoint *p = make_point(Lp[i].x, Lp[i].y, Lp[i].z, kframe, i, 1);
        p->mfMinDistance = dist_range[2 * i]; p->mfMaxDistance = dist_range[2 * i + 1];
        owned.push_back(p);
        kframe->mvpMapPoints[i] = p;
        index_of[p] = i;
        if (!Lp[i].valid) { if (kind == 1) p->mbBad = true; else found.insert(p); continue; }
        const cv::Mat PO = p->GetWorldPos() - Ow;
        const float dist3D = cv::norm(PO);
        gate_out[i] = !(dist3D < p->GetMinDistanceInvariance() || dist3D > p->GetMaxDistanceInvariance());
        if (gate_out[i]) level_out[i] = p->PredictScale(dist3D, cur);
    }
    claim(cur, Cur, kframe, owned);
    KeyFrame *kf = new KeyFrame(*kframe, the_map(), NULL);
    ORBmatcher matcher(nnratio, check_ori != 0);
    const int n = matcher.SearchByProjection(*cur, kf, found, th, orb_dist);
    for (int k = 0; k < Cur->n; k++) {
        std::map<MapPoint *, int>::const_iterator it = index_of.find(cur->mvpMapPoints[k]);
        match[k] = it == index_of.end() ? -1 : it->second;
    }
    delete kf;
    for (size_t i = 0; i < owned.size(); i++) delete owned[i];
    delete kframe; delete cur;
    orbref_arena_release();
    return n;
}

// ORBextractor::operator() through whatever ORBextractor this library was linked with (the reference's, or the adapter's)
int orbmref_extract(const uint8_t *img, int w, int h, int stride, int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th,
                    void *kps, uint8_t *desc, int cap) {
    orbref_arena_retain();
    ORBextractor *e = new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
    std::vector<cv::KeyPoint> keys;
    cv::Mat d;
    (*e)(cv::Mat(h, w, CV_8U, const_cast<uint8_t *>(img), (size_t)stride), cv::Mat(), keys, d);
    const int n = (int)keys.size();
    if (n <= cap && n > 0) {
        std::memcpy(kps, keys.data(), sizeof(cv::KeyPoint) * n);
        for (int i = 0; i < n; i++) std::memcpy(desc + 32 * (size_t)i, d.ptr(i), 32);
    }
    delete e;
    orbref_arena_release();
    return n;
}

// Frame::ComputeBoW (Frame.cc:423-431: transform(vCurrentDesc, mBowVec, mFeatVec, 4)) of a frame holding these descriptors, through
// whatever Frame::ComputeBoW this library was linked with (the reference's, or adapter/Frame_orbx.cc on the device vocabulary)
void orbvref_compute_bow(void *h, const uint8_t *desc, int n, uint32_t *bow_id, double *bow_val, int32_t *n_bow, uint32_t *fv_id,
                         int32_t *fv_start, uint32_t *fv_feat, int32_t *n_fv) {
    orbref_arena_retain();
    Frame *f = new Frame();
    f->N = n;
    f->mDescriptors = cv::Mat(n, 32, CV_8U, const_cast<uint8_t *>(desc)).clone();
    f->mpORBvocabulary = static_cast<ORBVocabulary *>(h);
    f->ComputeBoW();
    int k = 0, at = 0;
    for (DBoW2::BowVector::const_iterator it = f->mBowVec.begin(); it != f->mBowVec.end(); ++it, ++k) { bow_id[k] = it->first; bow_val[k] = it->second; }
    *n_bow = k;
    k = 0;
    for (DBoW2::FeatureVector::const_iterator it = f->mFeatVec.begin(); it != f->mFeatVec.end(); ++it, ++k) {
        fv_id[k] = it->first; fv_start[k] = at;
        for (size_t j = 0; j < it->second.size(); j++) fv_feat[at++] = it->second[j];
    }
    fv_start[k] = at;
    *n_fv = k;
    delete f;
    orbref_arena_release();
}

// ---- the window searches that mutate the map: Fuse (:825-975), Fuse with a Sim3 (:977-1100), SearchByProjection(KF, Scw, ...) (:290-403)
// A keyframe built by the reference's constructor from `KFrec` at pose (R, t); kf_has[k] != 0: keypoint k holds a map point that
// this keyframe (and kf_extra[k] further keyframes) observes.  Candidate i: pts[i] (position, normal, distance range; `skip` = 0
// good, 1 NULL, 2 bad, 3 already observed by the keyframe; `blocks` = number of other keyframes observing it) with descriptor row i.
// which = 0 Fuse(pKF, vpMapPoints, th); 1 Fuse(pKF, Scw, vpPoints, th, vpReplacePoint); 2 SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)
// with vpMatched preset to the keyframe's own points where kf_has[k] == 2.  Scw = [s R | s t].
// A MapPoint* is reported as -1 (NULL), i (candidate i) or 1000000 + k (the point keypoint k held at the start).
// Outputs: kf_slot[k] = pKF->GetMapPoint(k) afterwards; pt_bad / pt_obs / pt_replaced per candidate; kfmp_bad / kfmp_obs / kfmp_replaced per
// keypoint's original point; aux = vpReplacePoint (which 1, per candidate) or vpMatched (which 2, per keypoint).
int orbmref_window(int which, const orbo_frame *KFrec, const float R[9], const float t[3], float scale, const uint8_t *kf_has,
                   const int32_t *kf_extra, int n_pts, const orbo_frustum_point *pts, const uint8_t *pt_desc, float th, int32_t *kf_slot,
                   uint8_t *pt_bad, int32_t *pt_obs, int32_t *pt_replaced, uint8_t *kfmp_bad, int32_t *kfmp_obs, int32_t *kfmp_replaced,
                   int32_t *aux) {
    orbref_arena_retain();
    const int n = KFrec->n;
    Frame *f = make_frame(KFrec);
    f->SetPose(pose4(R, t));
    f->mvDepth.assign(n, -1.f);
    KeyFrame *kf = new KeyFrame(*f, the_map(), NULL);
    std::vector<KeyFrame *> others;
    for (int j = 0; j < 4; j++) others.push_back(new KeyFrame(*f, the_map(), NULL));     // further observers (same descriptors)
    orbo_frame C = *KFrec;
    std::vector<orbo_keypoint> ckeys(n_pts > 0 ? n_pts : 1);
    std::memset(ckeys.data(), 0, sizeof(orbo_keypoint) * ckeys.size());
    C.n = n_pts; C.keys_un = ckeys.data(); C.desc = pt_desc; C.u_right = NULL; C.claimed = NULL;
    Frame *carrier = make_frame(&C);
    std::vector<MapPoint *> kfmp(n, static_cast<MapPoint *>(NULL)), cand(n_pts, static_cast<MapPoint *>(NULL));
    std::map<MapPoint *, int> code;
    for (int k = 0; k < n; k++)
        if (kf_has[k]) {
            MapPoint *p = make_point(0, 0, 1 + k, f, k, 0);
            p->AddObservation(kf, k);
            kf->AddMapPoint(p, k);
            for (int j = 0; j < kf_extra[k] && j < 4; j++) { p->AddObservation(others[j], k); others[j]->AddMapPoint(p, k); }
            kfmp[k] = p;
            code[p] = 1000000 + k;
        }
    int free_slot = 0;
    for (int i = 0; i < n_pts; i++) {
        if (pts[i].skip == 1) continue;                      // a NULL entry of vpMapPoints
        MapPoint *p = make_point(pts[i].x, pts[i].y, pts[i].z, carrier, i, 0);
        p->mNormalVector = (cv::Mat_<float>(3, 1) << pts[i].nx, pts[i].ny, pts[i].nz);
        p->mfMinDistance = pts[i].min_distance; p->mfMaxDistance = pts[i].max_distance;
        for (int j = 0; j < pts[i].blocks && j < 4; j++) {   // seen from other keyframes, at keypoints those do not use otherwise
            while (free_slot < n && kf_has[free_slot]) free_slot++;
            if (free_slot >= n) break;
            p->AddObservation(others[j], free_slot);
        }
        if (pts[i].blocks) free_slot++;
        if (pts[i].skip == 3) {                              // already observed by this keyframe (IsInKeyFrame)
            while (free_slot < n && kf_has[free_slot]) free_slot++;
            if (free_slot < n) { p->AddObservation(kf, free_slot); kf->AddMapPoint(p, free_slot); free_slot++; }
        }
        if (pts[i].skip == 2) p->mbBad = true;
        cand[i] = p;
        code[p] = i;
    }
    struct Code { std::map<MapPoint *, int> &c; int operator()(MapPoint *p) const { if (!p) return -1; std::map<MapPoint *, int>::const_iterator it = c.find(p); return it == c.end() ? -2 : it->second; } } of = {code};
    cv::Mat Scw = cv::Mat::eye(4, 4, CV_32F);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Scw.at<float>(i, j) = scale * R[3 * i + j]; Scw.at<float>(i, 3) = scale * t[i]; }
    ORBmatcher matcher(0.8f, true);
    int ret = 0;
    if (which == 0) {
        ret = matcher.Fuse(kf, cand, th);
    } else if (which == 1) {
        std::vector<MapPoint *> good, replace;
        std::vector<int> idx;
        for (int i = 0; i < n_pts; i++) if (cand[i]) { good.push_back(cand[i]); idx.push_back(i); }      // LoopClosing passes no NULLs here
        replace.assign(good.size(), static_cast<MapPoint *>(NULL));
        ret = matcher.Fuse(kf, Scw, good, th, replace);
        for (int i = 0; i < n_pts; i++) aux[i] = -1;
        for (size_t j = 0; j < good.size(); j++) aux[idx[j]] = of(replace[j]);
    } else {
        std::vector<MapPoint *> good, matched(n, static_cast<MapPoint *>(NULL));
        for (int i = 0; i < n_pts; i++) if (cand[i]) good.push_back(cand[i]);
        for (int k = 0; k < n; k++) if (kf_has[k] == 2) matched[k] = kfmp[k];
        ret = matcher.SearchByProjection(kf, Scw, good, matched, (int)th);
        for (int k = 0; k < n; k++) aux[k] = of(matched[k]);
    }
    for (int k = 0; k < n; k++) {
        kf_slot[k] = of(kf->GetMapPoint(k));
        kfmp_bad[k] = kfmp[k] ? kfmp[k]->isBad() : 0;
        kfmp_obs[k] = kfmp[k] ? kfmp[k]->Observations() : -1;
        kfmp_replaced[k] = kfmp[k] ? of(kfmp[k]->GetReplaced()) : -1;
    }
    for (int i = 0; i < n_pts; i++) {
        pt_bad[i] = cand[i] ? cand[i]->isBad() : 0;
        pt_obs[i] = cand[i] ? cand[i]->Observations() : -1;
        pt_replaced[i] = cand[i] ? of(cand[i]->GetReplaced()) : -1;
    }
    for (int k = 0; k < n; k++) delete kfmp[k];
    for (int i = 0; i < n_pts; i++) delete cand[i];
    for (size_t j = 0; j < others.size(); j++) delete others[j];
    delete kf; delete carrier; delete f;
    orbref_arena_release();
    return ret;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th), ORBmatcher.cc:1102-1326 (loop closing).  Two keyframes built by the
// reference's constructor at poses (R1, t1), (R2, t2); keypoint k of keyframe j holds a map point where pj[k].skip == 0 (position
// and distance range from pj[k]; skip == 2: the point is bad).  preset12[i1] >= 0: vpMatches12[i1] starts as keyframe 2's point at that
// keypoint.  match12[i1] = keypoint of keyframe 2 whose point vpMatches12[i1] is afterwards, or -1.
int orbmref_search_by_sim3(const orbo_frame *K1, const float R1[9], const float t1[3], const orbo_frustum_point *p1, const orbo_frame *K2,
                           const float R2[9], const float t2[3], const orbo_frustum_point *p2, const int32_t *preset12, float s12,
                           const float R12[9], const float t12[3], float th, int32_t *match12) {
    orbref_arena_retain();
    Frame *f1 = make_frame(K1), *f2 = make_frame(K2);
    f1->SetPose(pose4(R1, t1));
    f2->SetPose(pose4(R2, t2));
    KeyFrame *k1 = new KeyFrame(*f1, the_map(), NULL), *k2 = new KeyFrame(*f2, the_map(), NULL);
    std::vector<MapPoint *> m1(K1->n, static_cast<MapPoint *>(NULL)), m2(K2->n, static_cast<MapPoint *>(NULL));
    std::map<MapPoint *, int> in2;
    for (int side = 0; side < 2; side++) {
        const orbo_frame *K = side ? K2 : K1;
        const orbo_frustum_point *P = side ? p2 : p1;
        Frame *f = side ? f2 : f1;
        KeyFrame *kf = side ? k2 : k1;
        std::vector<MapPoint *> &m = side ? m2 : m1;
        for (int k = 0; k < K->n; k++) {
            if (P[k].skip == 1) continue;
            MapPoint *p = make_point(P[k].x, P[k].y, P[k].z, f, k, 0);
            p->mfMinDistance = P[k].min_distance; p->mfMaxDistance = P[k].max_distance;
            p->AddObservation(kf, k);
            kf->AddMapPoint(p, k);
            if (P[k].skip == 2) p->mbBad = true;
            m[k] = p;
            if (side) in2[p] = k;
        }
    }
    std::vector<MapPoint *> matches(K1->n, static_cast<MapPoint *>(NULL));
    for (int i = 0; i < K1->n; i++)
        if (preset12[i] >= 0 && preset12[i] < K2->n) matches[i] = m2[preset12[i]];
    cv::Mat R(3, 3, CV_32F), t(3, 1, CV_32F);
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R.at<float>(i, j) = R12[3 * i + j]; t.at<float>(i) = t12[i]; }
    ORBmatcher matcher(0.75f, true);
    const int n = matcher.SearchBySim3(k1, k2, matches, s12, R, t, th);
    for (int i = 0; i < K1->n; i++) {
        std::map<MapPoint *, int>::const_iterator it = in2.find(matches[i]);
        match12[i] = it == in2.end() ? -1 : it->second;
    }
    for (int k = 0; k < K1->n; k++) delete m1[k];
    for (int k = 0; k < K2->n; k++) delete m2[k];
    delete k1; delete k2; delete f1; delete f2;
    orbref_arena_release();
    return n;
}

// ---- small members ----
// ORBmatcher::ComputeThreeMaxima (ORBmatcher.cc:1601-1642; protected): histo[k] = number of entries of bin k.  Reference build only:
// the adapters run the histogram on the device and do not define this helper.
#ifndef ORBREF_SYSTEM_ALLOCATOR
void orbmref_three_maxima(const int32_t *histo, int L, int32_t *ind1, int32_t *ind2, int32_t *ind3) {
    orbref_arena_retain();
    {
        std::vector<std::vector<int> > h(L);
        for (int k = 0; k < L; k++) h[k].assign(histo[k], 0);
        ORBmatcher m(0.6f, true);
        int a = -1, b = -1, c = -1;
        m.ComputeThreeMaxima(h.data(), L, a, b, c);
        *ind1 = a; *ind2 = b; *ind3 = c;
    }
    orbref_arena_release();
}
#endif
// KeyFrame::GetFeaturesInArea(x, y, r) (KeyFrame.cc:630-669) on a keyframe built from the record
int orbmref_keyframe_features_in_area(const orbo_frame *F, float x, float y, float r, int *out) {
    orbref_arena_retain();
    Frame *f = make_frame(F);
    KeyFrame *kf = new KeyFrame(*f, the_map(), NULL);
    const std::vector<size_t> v = kf->GetFeaturesInArea(x, y, r);
    for (size_t i = 0; i < v.size(); i++) out[i] = (int)v[i];
    const int n = (int)v.size();
    delete kf; delete f;
    orbref_arena_release();
    return n;
}
// the protected constructor tables of ORBextractor (ORBextractor.cc:436-469): features per level and umax[0..15]
void orbmref_extractor_quota_umax(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int32_t *quota, int32_t *umax16) {
    orbref_arena_retain();
    ORBextractor *e = new ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
    for (int l = 0; l < nlevels; l++) quota[l] = e->mnFeaturesPerLevel[l];
    for (int v = 0; v < 16; v++) umax16[v] = v < (int)e->umax.size() ? e->umax[v] : -1;
    delete e;
    orbref_arena_release();
}
}
