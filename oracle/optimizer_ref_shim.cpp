// C entry points around the REFERENCE's own Optimizer::LocalBundleAdjustment / Optimizer::PoseOptimization (src/Optimizer.cc),
// Converter (src/Converter.cc) and the vendored g2o (Thirdparty/g2o/g2o/{core,types,solvers,stuff}), all compiled UNMODIFIED from
// where they lie under /root/reference against the two stand-ins oracle/cvmini (OpenCV) and oracle/eigenmini (Eigen).
// TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref_optimizer.py drives it to pin oracle/lba_oracle.c and oracle/pose_oracle.c
// against the reference's code, and tests/test_adapter_optimizer_gpu.py runs the same entry points once against this library
// (the reference's Optimizer) and once against the drop-in build (adapter/Optimizer_orbx.cc on liborbx.so).
//
// Two kinds of entry points:
//   * leaves: one g2o edge / vertex / kernel / Converter function evaluated on given numbers (optref_edge_*, optref_se3_oplus,
//     optref_huber, optref_to_se3quat, optref_to_cvmat);
//   * whole functions: a KeyFrame / MapPoint / Map graph (or a Frame) is built from plain arrays with the reference's own
//     constructors, AddObservation / AddMapPoint / UpdateConnections, and the reference's member is called on it
//     (optref_local_ba, optref_pose_optimization).
#include <opencv2/core/core.hpp>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include "Optimizer.h"
#include "Converter.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "KeyFrameDatabase.h"
#include "Map.h"
#include "MapPoint.h"
#include "Thirdparty/g2o/g2o/core/jacobian_workspace.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"

namespace ORB_SLAM2 {
// src/Map.cc and src/KeyFrameDatabase.cc are not part of this build (this fork's Map.cc carries a FileStorage-based map loader).
// The optimisers only take Map::mMutexMapUpdate; MapPoint::SetBadFlag ends in mpMap->EraseMapPoint(this) (MapPoint.cc:189), which
// removes the point from the map's own set.
Map::Map() : mnMaxKFid(0), mnBigChangeIdx(0) {}
void Map::AddKeyFrame(KeyFrame *p) { std::unique_lock<std::mutex> lock(mMutexMap); mspKeyFrames.insert(p); if (p->mnId > mnMaxKFid) mnMaxKFid = p->mnId; }
void Map::AddMapPoint(MapPoint *p) { std::unique_lock<std::mutex> lock(mMutexMap); mspMapPoints.insert(p); }
void Map::EraseMapPoint(MapPoint *p) { std::unique_lock<std::mutex> lock(mMutexMap); mspMapPoints.erase(p); }
void Map::EraseKeyFrame(KeyFrame *p) { std::unique_lock<std::mutex> lock(mMutexMap); mspKeyFrames.erase(p); }
std::vector<KeyFrame *> Map::GetAllKeyFrames() { std::unique_lock<std::mutex> lock(mMutexMap); return std::vector<KeyFrame *>(mspKeyFrames.begin(), mspKeyFrames.end()); }
std::vector<MapPoint *> Map::GetAllMapPoints() { std::unique_lock<std::mutex> lock(mMutexMap); return std::vector<MapPoint *>(mspMapPoints.begin(), mspMapPoints.end()); }
long unsigned int Map::GetMaxKFid() { std::unique_lock<std::mutex> lock(mMutexMap); return mnMaxKFid; }
void KeyFrameDatabase::erase(KeyFrame *) {}
}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

namespace {
cv::Mat mat4(const float *T) {
    cv::Mat M(4, 4, CV_32F);
    for (int i = 0; i < 16; i++) M.at<float>(i / 4, i % 4) = T[i];
    return M;
}
g2o::SE3Quat se3_of(const double p[7]) {            // the given coefficients, bit for bit (the (q, t) constructor would re-normalise)
    g2o::SE3Quat T;
    T.setRotation(Eigen::Quaterniond(p[3], p[0], p[1], p[2]));
    T.setTranslation(Eigen::Vector3d(p[4], p[5], p[6]));
    return T;
}
void se3_out(const g2o::SE3Quat &T, double p[7]) {
    p[0] = T.rotation().x(); p[1] = T.rotation().y(); p[2] = T.rotation().z(); p[3] = T.rotation().w();
    for (int i = 0; i < 3; i++) p[4 + i] = T.translation()[i];
}
template <class M> void mat_out(const M &m, double *o) {   // row-major
    for (int r = 0; r < m.rows(); r++) for (int c = 0; c < m.cols(); c++) o[r * m.cols() + c] = m(r, c);
}

struct Cam { float fx, fy, cx, cy, bf; int nlevels; float scale_factor; };

// a Frame as Frame.cc's constructors leave the members that KeyFrame's constructor and the optimisers read
Frame *make_frame(const Cam &K, int n, const float *xy_ur, const int32_t *octave, const float *Tcw) {
    Frame *f = new Frame();
    f->N = n;
    f->mvKeys.resize(n);
    for (int i = 0; i < n; i++) { f->mvKeys[i] = cv::KeyPoint(xy_ur[3 * i], xy_ur[3 * i + 1], 31.f, 0.f, 20.f, octave[i]); }
    f->mvKeysUn = f->mvKeys;
    f->mDescriptors = cv::Mat(std::max(n, 1), 32, CV_8U);
    std::memset(f->mDescriptors.data, 0, (size_t)std::max(n, 1) * 32);
    f->mvuRight.resize(n); f->mvDepth.resize(n);
    for (int i = 0; i < n; i++) { f->mvuRight[i] = xy_ur[3 * i + 2]; f->mvDepth[i] = xy_ur[3 * i + 2] < 0 ? -1.f : K.bf / std::max(xy_ur[3 * i] - xy_ur[3 * i + 2], 1e-3f); }
    f->mvpMapPoints.assign(n, static_cast<MapPoint *>(NULL));
    f->mvbOutlier.assign(n, false);
    Frame::fx = K.fx; Frame::fy = K.fy; Frame::cx = K.cx; Frame::cy = K.cy; Frame::invfx = 1.0f / K.fx; Frame::invfy = 1.0f / K.fy;
    f->mbf = K.bf; f->mb = K.bf / K.fx;
    f->mK = cv::Mat::eye(3, 3, CV_32F);
    f->mK.at<float>(0, 0) = K.fx; f->mK.at<float>(1, 1) = K.fy; f->mK.at<float>(0, 2) = K.cx; f->mK.at<float>(1, 2) = K.cy;
    Frame::mnMinX = 0; Frame::mnMinY = 0; Frame::mnMaxX = 2 * K.cx; Frame::mnMaxY = 2 * K.cy;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / (Frame::mnMaxX - Frame::mnMinX);
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / (Frame::mnMaxY - Frame::mnMinY);
    Frame::mbInitialComputations = false;
    f->mnScaleLevels = K.nlevels;
    f->mfScaleFactor = K.scale_factor;
    f->mfLogScaleFactor = std::log(K.scale_factor);
    f->mvScaleFactors.resize(K.nlevels); f->mvInvScaleFactors.resize(K.nlevels); f->mvLevelSigma2.resize(K.nlevels); f->mvInvLevelSigma2.resize(K.nlevels);
    f->mvScaleFactors[0] = 1.0f; f->mvLevelSigma2[0] = 1.0f;                       // ORBextractor.cc:417-431
    for (int l = 1; l < K.nlevels; l++) { f->mvScaleFactors[l] = f->mvScaleFactors[l - 1] * K.scale_factor; f->mvLevelSigma2[l] = f->mvScaleFactors[l] * f->mvScaleFactors[l]; }
    for (int l = 0; l < K.nlevels; l++) { f->mvInvScaleFactors[l] = 1.0f / f->mvScaleFactors[l]; f->mvInvLevelSigma2[l] = 1.0f / f->mvLevelSigma2[l]; }
    f->mnId = Frame::nNextId++;
    f->SetPose(mat4(Tcw));
    return f;
}
}  // namespace

extern "C" {

// ------------------------------------------------------------------ leaves ------------------------------------------------
// EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ (types_six_dof_expmap.h:77-127, .cpp:103-234): error, chi2 with information
// inv_sigma2 * I, isDepthPositive, and the two Jacobian blocks (row-major: dE/dX is D x 3, dE/dxi is D x 6)
void optref_edge_binary(int stereo, const double pose[7], const double X[3], const double obs[3], const double K[5], float inv_sigma2,
                        double *err, double *chi2, int *depth_positive, double *JX, double *Jxi) {
    g2o::VertexSBAPointXYZ vp;
    vp.setEstimate(Eigen::Vector3d(X[0], X[1], X[2]));
    vp.setId(1);
    g2o::VertexSE3Expmap vc;
    vc.setEstimate(se3_of(pose));
    vc.setId(0);
    g2o::JacobianWorkspace jw;
    if (!stereo) {
        g2o::EdgeSE3ProjectXYZ e;
        e.setVertex(0, &vp); e.setVertex(1, &vc);
        Eigen::Matrix<double, 2, 1> o; o << obs[0], obs[1];
        e.setMeasurement(o);
        e.setInformation(Eigen::Matrix2d::Identity() * inv_sigma2);
        e.fx = K[0]; e.fy = K[1]; e.cx = K[2]; e.cy = K[3];
        e.computeError();
        jw.updateSize(&e); jw.allocate();
        static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(jw);
        for (int i = 0; i < 2; i++) err[i] = e.error()[i];
        *chi2 = e.chi2(); *depth_positive = e.isDepthPositive();
        mat_out(e.jacobianOplusXi(), JX); mat_out(e.jacobianOplusXj(), Jxi);
    } else {
        g2o::EdgeStereoSE3ProjectXYZ e;
        e.setVertex(0, &vp); e.setVertex(1, &vc);
        Eigen::Matrix<double, 3, 1> o; o << obs[0], obs[1], obs[2];
        e.setMeasurement(o);
        e.setInformation(Eigen::Matrix3d::Identity() * inv_sigma2);
        e.fx = K[0]; e.fy = K[1]; e.cx = K[2]; e.cy = K[3]; e.bf = K[4];
        e.computeError();
        jw.updateSize(&e); jw.allocate();
        static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(jw);
        for (int i = 0; i < 3; i++) err[i] = e.error()[i];
        *chi2 = e.chi2(); *depth_positive = e.isDepthPositive();
        mat_out(e.jacobianOplusXi(), JX); mat_out(e.jacobianOplusXj(), Jxi);
    }
}
// EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose (types_six_dof_expmap.h:130-208, .cpp:266-364)
void optref_edge_pose_only(int stereo, const double pose[7], const double X[3], const double obs[3], const double K[5], float inv_sigma2,
                           double *err, double *chi2, int *depth_positive, double *Jxi) {
    g2o::VertexSE3Expmap vc;
    vc.setEstimate(se3_of(pose));
    vc.setId(0);
    g2o::JacobianWorkspace jw;
    if (!stereo) {
        g2o::EdgeSE3ProjectXYZOnlyPose e;
        e.setVertex(0, &vc);
        Eigen::Matrix<double, 2, 1> o; o << obs[0], obs[1];
        e.setMeasurement(o);
        e.setInformation(Eigen::Matrix2d::Identity() * inv_sigma2);
        e.fx = K[0]; e.fy = K[1]; e.cx = K[2]; e.cy = K[3];
        e.Xw = Eigen::Vector3d(X[0], X[1], X[2]);
        e.computeError();
        jw.updateSize(&e); jw.allocate();
        static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(jw);
        for (int i = 0; i < 2; i++) err[i] = e.error()[i];
        *chi2 = e.chi2(); *depth_positive = e.isDepthPositive();
        mat_out(e.jacobianOplusXi(), Jxi);
    } else {
        g2o::EdgeStereoSE3ProjectXYZOnlyPose e;
        e.setVertex(0, &vc);
        Eigen::Matrix<double, 3, 1> o; o << obs[0], obs[1], obs[2];
        e.setMeasurement(o);
        e.setInformation(Eigen::Matrix3d::Identity() * inv_sigma2);
        e.fx = K[0]; e.fy = K[1]; e.cx = K[2]; e.cy = K[3]; e.bf = K[4];
        e.Xw = Eigen::Vector3d(X[0], X[1], X[2]);
        e.computeError();
        jw.updateSize(&e); jw.allocate();
        static_cast<g2o::OptimizableGraph::Edge &>(e).linearizeOplus(jw);
        for (int i = 0; i < 3; i++) err[i] = e.error()[i];
        *chi2 = e.chi2(); *depth_positive = e.isDepthPositive();
        mat_out(e.jacobianOplusXi(), Jxi);
    }
}
// VertexSE3Expmap::oplusImpl = SE3Quat::exp(update) * estimate (types_six_dof_expmap.h:73-76, se3quat.h:223-257)
void optref_se3_oplus(const double pose[7], const double update[6], double out[7]) {
    g2o::VertexSE3Expmap v;
    v.setEstimate(se3_of(pose));
    v.oplus(update);
    se3_out(v.estimate(), out);
}
// SE3Quat::map (se3quat.h:217-220)
void optref_se3_map(const double pose[7], const double X[3], double out[3]) {
    const Eigen::Vector3d y = se3_of(pose).map(Eigen::Vector3d(X[0], X[1], X[2]));
    for (int i = 0; i < 3; i++) out[i] = y[i];
}
// RobustKernelHuber::robustify (robust_kernel_impl.cpp:78-91)
void optref_huber(double e2, double delta, double rho[3]) {
    g2o::RobustKernelHuber rk;
    rk.setDelta(delta);
    Eigen::Vector3d r;
    rk.robustify(e2, r);
    for (int i = 0; i < 3; i++) rho[i] = r[i];
}
// Converter::toSE3Quat(cv::Mat) and Converter::toCvMat(SE3Quat) (Converter.cc:41-53, 63-67)
void optref_to_se3quat(const float Tcw[16], double pose[7]) { se3_out(Converter::toSE3Quat(mat4(Tcw)), pose); }
void optref_to_cvmat(const double pose[7], float Tcw[16]) {
    const cv::Mat M = Converter::toCvMat(se3_of(pose));
    for (int i = 0; i < 16; i++) Tcw[i] = M.at<float>(i / 4, i % 4);
}

// ------------------------------------------------------------------ whole functions ---------------------------------------
typedef struct {
    int32_t n_kf;
    const float *kf_Tcw;           /* n_kf x 16, row-major Tcw */
    const int32_t *kf_start;       /* n_kf + 1: keypoints of keyframe j are rows kf_start[j] .. kf_start[j+1] of the kp_* arrays */
    const float *kp_xy_ur;         /* x, y, uRight (< 0 = monocular) */
    const int32_t *kp_octave;
    const int32_t *kp_point;       /* map point index or -1 */
    int32_t n_pts;
    const float *pts;              /* n_pts x 3 world positions */
    float fx, fy, cx, cy, bf;
    int32_t nlevels;
    float scale_factor;
    int32_t center_kf;             /* pKF */
    int32_t first_kf_id;           /* mnId of keyframe 0 (KeyFrame::nNextId is set to it) */
    int32_t stop_before;           /* *pbStopFlag = true on entry */
} optref_lba_graph;

/* out_Tcw n_kf x 16, out_pts n_pts x 3, out_kp_kept per keypoint (1 = still holds its map point), out_pt_bad per point,
 * out_kf_role per keyframe: 1 = local (free unless mnId == 0), 2 = fixed, 0 = not in the window */
int optref_local_ba(const optref_lba_graph *G, float *out_Tcw, float *out_pts, uint8_t *out_kp_kept, uint8_t *out_pt_bad,
                    uint8_t *out_kf_role) {
    const Cam K = {G->fx, G->fy, G->cx, G->cy, G->bf, G->nlevels, G->scale_factor};
    Map *map = new Map();
    KeyFrame::nNextId = G->first_kf_id;
    std::vector<Frame *> frames;
    std::vector<KeyFrame *> kfs;
    for (int j = 0; j < G->n_kf; j++) {
        const int s = G->kf_start[j], n = G->kf_start[j + 1] - s;
        frames.push_back(make_frame(K, n, G->kp_xy_ur + 3 * s, G->kp_octave + s, G->kf_Tcw + 16 * j));
        kfs.push_back(new KeyFrame(*frames.back(), map, NULL));
        map->AddKeyFrame(kfs.back());
    }
    // every map point is created from its first observing keyframe (MapPoint.cc:58-72), then observed (MapPoint.cc:102-115)
    std::vector<MapPoint *> mps(G->n_pts, static_cast<MapPoint *>(NULL));
    for (int j = 0; j < G->n_kf; j++)
        for (int k = G->kf_start[j]; k < G->kf_start[j + 1]; k++) {
            const int p = G->kp_point[k];
            if (p < 0) continue;
            if (!mps[p]) {
                cv::Mat P(3, 1, CV_32F);
                for (int i = 0; i < 3; i++) P.at<float>(i) = G->pts[3 * p + i];
                mps[p] = new MapPoint(P, kfs[j], map);
                map->AddMapPoint(mps[p]);
            }
            mps[p]->AddObservation(kfs[j], k - G->kf_start[j]);
            kfs[j]->AddMapPoint(mps[p], k - G->kf_start[j]);
        }
    for (int j = 0; j < G->n_kf; j++) kfs[j]->UpdateConnections();
    for (int p = 0; p < G->n_pts; p++) if (mps[p]) mps[p]->UpdateNormalAndDepth();

    bool stop = G->stop_before != 0;
    Optimizer::LocalBundleAdjustment(kfs[G->center_kf], &stop, map);

    const unsigned long id = kfs[G->center_kf]->mnId;
    for (int j = 0; j < G->n_kf; j++) {
        const cv::Mat T = kfs[j]->GetPose();
        for (int i = 0; i < 16; i++) out_Tcw[16 * j + i] = T.at<float>(i / 4, i % 4);
        out_kf_role[j] = kfs[j]->mnBALocalForKF == id ? 1 : (kfs[j]->mnBAFixedForKF == id ? 2 : 0);
        for (int k = G->kf_start[j]; k < G->kf_start[j + 1]; k++) {
            const int p = G->kp_point[k];
            out_kp_kept[k] = p >= 0 && kfs[j]->GetMapPoint(k - G->kf_start[j]) == mps[p] && mps[p]->IsInKeyFrame(kfs[j]);
        }
    }
    for (int p = 0; p < G->n_pts; p++) {
        if (!mps[p]) { out_pt_bad[p] = 2; for (int i = 0; i < 3; i++) out_pts[3 * p + i] = G->pts[3 * p + i]; continue; }
        const cv::Mat X = mps[p]->GetWorldPos();
        for (int i = 0; i < 3; i++) out_pts[3 * p + i] = X.at<float>(i);
        out_pt_bad[p] = mps[p]->isBad();
    }
    for (size_t i = 0; i < mps.size(); i++) delete mps[i];
    for (size_t i = 0; i < kfs.size(); i++) { delete kfs[i]; delete frames[i]; }
    delete map;
    return 0;
}

typedef struct {
    int32_t n;                     /* keypoints of the frame */
    const float *kp_xy_ur;         /* x, y, uRight (< 0 = monocular) */
    const int32_t *kp_octave;
    const float *Xw;               /* n x 3: the map point matched to each keypoint */
    const uint8_t *has_point;      /* mvpMapPoints[i] != NULL */
    float Tcw[16];
    float fx, fy, cx, cy, bf;
    int32_t nlevels;
    float scale_factor;
} optref_pose_frame;

/* returns Optimizer::PoseOptimization(&frame); out_Tcw = pFrame->mTcw afterwards, outlier = mvbOutlier, n_bad = nBadPoseOpt */
int optref_pose_optimization(const optref_pose_frame *F, float out_Tcw[16], uint8_t *outlier, int32_t *n_bad) {
    const Cam K = {F->fx, F->fy, F->cx, F->cy, F->bf, F->nlevels, F->scale_factor};
    Map *map = new Map();
    Frame *f = make_frame(K, F->n, F->kp_xy_ur, F->kp_octave, F->Tcw);
    std::vector<MapPoint *> mps;
    for (int i = 0; i < F->n; i++) {
        if (!F->has_point[i]) continue;
        cv::Mat P(3, 1, CV_32F);
        for (int k = 0; k < 3; k++) P.at<float>(k) = F->Xw[3 * i + k];
        mps.push_back(new MapPoint(P, map, f, i));
        f->mvpMapPoints[i] = mps.back();
    }
    f->nBadPoseOpt = 0;
    const int r = Optimizer::PoseOptimization(f);
    for (int i = 0; i < 16; i++) out_Tcw[i] = f->mTcw.at<float>(i / 4, i % 4);
    for (int i = 0; i < F->n; i++) outlier[i] = f->mvbOutlier[i];
    *n_bad = f->nBadPoseOpt;
    for (size_t i = 0; i < mps.size(); i++) delete mps[i];
    delete f;
    delete map;
    return r;
}

}  // extern "C"
