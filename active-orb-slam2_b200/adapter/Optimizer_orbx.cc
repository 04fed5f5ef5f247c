// ORB_SLAM2::Optimizer::LocalBundleAdjustment and Optimizer::PoseOptimization on top of the orbx C ABI (include/orbx.h).
// include/Optimizer.h stays as it is; the other members of the class (BundleAdjustment, GlobalBundleAdjustemnt,
// OptimizeEssentialGraph, OptimizeSim3) stay in the reference's src/Optimizer.cc.
//
// Build: compile this file next to src/Optimizer.cc with the definitions of these two members removed there (or renamed, as the
// drop-in build in oracle/Makefile does with -DLocalBundleAdjustment=LocalBundleAdjustment_reference).  Callers link unchanged:
// LocalMapping.cc:81 (LocalBundleAdjustment), Tracking.cc:870, :994, :1039, :1620 (PoseOptimization).
//
// What stays on the host, in the reference's own order and under the reference's own locks: collecting the window
// (Optimizer.cc:456-505: covisible keyframes, their map points, the fixed observers, with the same mnBALocalForKF / mnBAFixedForKF
// marks), the boundary casts (Converter::toSE3Quat / toVector3d / toCvMat, Optimizer.cc:528, :573, :766, :774), the erase list and
// the write-back under Map::mMutexMapUpdate (:745-778).  What moves to the device: everything g2o did in between (:659-735).
//
// Error behaviour: the reference has none (void / count returns, no exceptions).  A device failure is reported on stderr and the
// call returns without touching the map (LocalBundleAdjustment) or with the frame's pose unchanged (PoseOptimization returns 0,
// which Tracking treats as "lost": the same path as a failed optimisation); nothing is thrown into Tracking / LocalMapping.
#include "Optimizer.h"

#include "Converter.h"

#include "orbx_adapter.h"

#include <algorithm>
#include <list>
#include <mutex>
#include <vector>

using namespace std;

namespace ORB_SLAM2
{

namespace
{
// LocalBundleAdjustment runs on the LocalMapping thread, PoseOptimization on the Tracking thread: one handle per calling thread,
// sized by the largest window / frame seen so far (the reference allocates a fresh g2o graph per call)
struct LbaHandle
{
    orbx_lba* h = nullptr;
    int kf = 0, pts = 0, edges = 0;
    ~LbaHandle() { if (h) orbx_lba_destroy(h); }
    orbx_lba* fit(int nKF, int nPts, int nEdges)
    {
        if (h && nKF <= kf && nPts <= pts && nEdges <= edges)
            return h;
        if (h) { orbx_lba_destroy(h); h = nullptr; }
        kf = std::max(64, nKF + nKF / 2); pts = std::max(8192, nPts + nPts / 2); edges = std::max(65536, nEdges + nEdges / 2);
        if (orbxFailed(orbx_lba_create(&h, kf, pts, edges, orbxDevice()), "orbx_lba_create"))
        {
            h = nullptr;
            kf = pts = edges = 0;
        }
        return h;
    }
};
struct PoseHandle
{
    orbx_pose* h = nullptr;
    int obs = 0;
    ~PoseHandle() { if (h) orbx_pose_destroy(h); }
    orbx_pose* fit(int nObs)
    {
        if (h && nObs <= obs)
            return h;
        if (h) { orbx_pose_destroy(h); h = nullptr; }
        obs = std::max(8192, nObs + nObs / 2);
        if (orbxFailed(orbx_pose_create(&h, obs, 1, orbxDevice()), "orbx_pose_create"))
        {
            h = nullptr;
            obs = 0;
        }
        return h;
    }
};

void poseOf(const g2o::SE3Quat& T, double p[7])
{
    p[0] = T.rotation().x(); p[1] = T.rotation().y(); p[2] = T.rotation().z(); p[3] = T.rotation().w();
    for (int i = 0; i < 3; i++) p[4 + i] = T.translation()[i];
}
g2o::SE3Quat se3Of(const double p[7])
{
    g2o::SE3Quat T;                                    // the optimiser's coefficients as they are (the (q, t) constructor re-normalises)
    T.setRotation(Eigen::Quaterniond(p[3], p[0], p[1], p[2]));
    T.setTranslation(Eigen::Vector3d(p[4], p[5], p[6]));
    return T;
}
} // namespace

// replaces Optimizer.cc:454-779
void Optimizer::LocalBundleAdjustment(KeyFrame* pKF, bool* pbStopFlag, Map* pMap)
{
    // Local KeyFrames: first breadth search from the current keyframe (:456-470)
    list<KeyFrame*> lLocalKeyFrames;
    lLocalKeyFrames.push_back(pKF);
    pKF->mnBALocalForKF = pKF->mnId;
    const vector<KeyFrame*> vNeighKFs = pKF->GetVectorCovisibleKeyFrames();
    for (int i = 0, iend = vNeighKFs.size(); i < iend; i++)
    {
        KeyFrame* pKFi = vNeighKFs[i];
        pKFi->mnBALocalForKF = pKF->mnId;
        if (!pKFi->isBad())
            lLocalKeyFrames.push_back(pKFi);
    }
    // Local MapPoints seen in local keyframes (:472-488)
    list<MapPoint*> lLocalMapPoints;
    for (list<KeyFrame*>::iterator lit = lLocalKeyFrames.begin(), lend = lLocalKeyFrames.end(); lit != lend; lit++)
    {
        vector<MapPoint*> vpMPs = (*lit)->GetMapPointMatches();
        for (vector<MapPoint*>::iterator vit = vpMPs.begin(), vend = vpMPs.end(); vit != vend; vit++)
        {
            MapPoint* pMP = *vit;
            if (pMP && !pMP->isBad() && pMP->mnBALocalForKF != pKF->mnId)
            {
                lLocalMapPoints.push_back(pMP);
                pMP->mnBALocalForKF = pKF->mnId;
            }
        }
    }
    // Fixed keyframes: they see local map points but are not local keyframes (:490-505)
    list<KeyFrame*> lFixedCameras;
    for (list<MapPoint*>::iterator lit = lLocalMapPoints.begin(), lend = lLocalMapPoints.end(); lit != lend; lit++)
    {
        map<KeyFrame*, size_t> observations = (*lit)->GetObservations();
        for (map<KeyFrame*, size_t>::iterator mit = observations.begin(), mend = observations.end(); mit != mend; mit++)
        {
            KeyFrame* pKFi = mit->first;
            if (pKFi->mnBALocalForKF != pKF->mnId && pKFi->mnBAFixedForKF != pKF->mnId)
            {
                pKFi->mnBAFixedForKF = pKF->mnId;
                if (!pKFi->isBad())
                    lFixedCameras.push_back(pKFi);
            }
        }
    }

    // keyframe vertices (:521-545): local ones first, then the fixed ones
    std::vector<KeyFrame*> vpKF;
    std::vector<double> kfPose;
    std::vector<uint8_t> kfFixed;
    std::map<KeyFrame*, int> kfIndex;
    for (int pass = 0; pass < 2; pass++)
    {
        list<KeyFrame*>& l = pass == 0 ? lLocalKeyFrames : lFixedCameras;
        for (list<KeyFrame*>::iterator lit = l.begin(), lend = l.end(); lit != lend; lit++)
        {
            KeyFrame* pKFi = *lit;
            double p[7];
            poseOf(Converter::toSE3Quat(pKFi->GetPose()), p);
            kfIndex[pKFi] = (int)vpKF.size();
            vpKF.push_back(pKFi);
            kfPose.insert(kfPose.end(), p, p + 7);
            kfFixed.push_back(pass == 1 || pKFi->mnId == 0);
        }
    }
    // map point vertices and one edge per observation (:571-651), in the reference's order
    std::vector<MapPoint*> vpMP;
    std::vector<double> pts, eObs;
    std::vector<int32_t> eKF, ePt;
    std::vector<float> eInvSigma2;
    std::vector<uint8_t> eStereo;
    std::vector<KeyFrame*> vpEdgeKF;
    std::vector<MapPoint*> vpEdgeMP;
    double K[5] = {0, 0, 0, 0, 0};
    for (list<MapPoint*>::iterator lit = lLocalMapPoints.begin(), lend = lLocalMapPoints.end(); lit != lend; lit++)
    {
        MapPoint* pMP = *lit;
        const Eigen::Matrix<double, 3, 1> X = Converter::toVector3d(pMP->GetWorldPos());
        const int ip = (int)vpMP.size();
        vpMP.push_back(pMP);
        for (int i = 0; i < 3; i++) pts.push_back(X[i]);
        const map<KeyFrame*, size_t> observations = pMP->GetObservations();
        for (map<KeyFrame*, size_t>::const_iterator mit = observations.begin(), mend = observations.end(); mit != mend; mit++)
        {
            KeyFrame* pKFi = mit->first;
            if (pKFi->isBad())
                continue;
            std::map<KeyFrame*, int>::const_iterator ik = kfIndex.find(pKFi);
            if (ik == kfIndex.end())
                continue;   // a keyframe that went bad between the two passes above; optimizer.vertex(id) would be NULL in the reference
            const cv::KeyPoint& kpUn = pKFi->mvKeysUn[mit->second];
            const float ur = pKFi->mvuRight[mit->second];
            eKF.push_back(ik->second); ePt.push_back(ip);
            eObs.push_back(kpUn.pt.x); eObs.push_back(kpUn.pt.y); eObs.push_back(ur);
            eInvSigma2.push_back(pKFi->mvInvLevelSigma2[kpUn.octave]);
            eStereo.push_back(!(ur < 0));                                        // :595
            vpEdgeKF.push_back(pKFi); vpEdgeMP.push_back(pMP);
            K[0] = pKFi->fx; K[1] = pKFi->fy; K[2] = pKFi->cx; K[3] = pKFi->cy; K[4] = pKFi->mbf;   // e->fx = pKFi->fx ... (:611-614, :640-644)
        }
    }

    if (pbStopFlag && *pbStopFlag)                                               // :653-655
        return;

    const int nKF = (int)vpKF.size(), nPts = (int)vpMP.size(), nEdges = (int)eKF.size();
    thread_local LbaHandle handle;
    orbx_lba* h = handle.fit(nKF, nPts, nEdges);
    if (!h)
        return;
    orbx_lba_problem P = orbx_lba_problem();
    P.n_kf = nKF; P.kf_pose = kfPose.data(); P.kf_fixed = kfFixed.data();
    P.n_pts = nPts; P.pts = pts.data();
    P.n_edges = nEdges; P.e_kf = eKF.data(); P.e_pt = ePt.data(); P.e_obs = eObs.data(); P.e_inv_sigma2 = eInvSigma2.data(); P.e_stereo = eStereo.data();
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3]; P.bf = K[4];
    static_assert(sizeof(bool) == 1, "pbStopFlag is polled as a byte");
    P.stop_flag = reinterpret_cast<const volatile uint8_t*>(pbStopFlag);
    std::vector<double> kfOut((size_t)7 * std::max(nKF, 1)), ptOut((size_t)3 * std::max(nPts, 1));
    std::vector<uint8_t> erase(std::max(nEdges, 1), 0);
    orbx_lba_result R = orbx_lba_result();
    R.kf_pose = kfOut.data(); R.pts = ptOut.data(); R.erase = erase.data();
    // optimize(5) with Huber kernels, classification, optimize(10) without them, final classification (:659-735)
    if (orbxFailed(orbx_lba_solve_host(h, &P, 5, 10, &R), "orbx_lba_solve_host") || R.stopped)
        return;

    // the erase list in the reference's order: monocular edges first, then stereo edges (:709-743)
    vector<pair<KeyFrame*, MapPoint*> > vToErase;
    vToErase.reserve(nEdges);
    for (int stereo = 0; stereo < 2; stereo++)
        for (int e = 0; e < nEdges; e++)
            if (eStereo[e] == stereo && !vpEdgeMP[e]->isBad() && erase[e])
                vToErase.push_back(make_pair(vpEdgeKF[e], vpEdgeMP[e]));

    unique_lock<mutex> lock(pMap->mMutexMapUpdate);                              // :746
    for (size_t i = 0; i < vToErase.size(); i++)
    {
        KeyFrame* pKFi = vToErase[i].first;
        MapPoint* pMPi = vToErase[i].second;
        pKFi->EraseMapPointMatch(pMPi);
        pMPi->EraseObservation(pKFi);
    }
    // recover the optimised data (:760-778): local keyframes, then every local map point
    for (int k = 0; k < (int)lLocalKeyFrames.size(); k++)
        vpKF[k]->SetPose(Converter::toCvMat(se3Of(&kfOut[(size_t)7 * k])));
    for (int l = 0; l < nPts; l++)
    {
        Eigen::Matrix<double, 3, 1> X;
        X << ptOut[(size_t)3 * l], ptOut[(size_t)3 * l + 1], ptOut[(size_t)3 * l + 2];
        vpMP[l]->SetWorldPos(Converter::toCvMat(X));
        vpMP[l]->UpdateNormalAndDepth();
    }
}

// replaces Optimizer.cc:239-452
int Optimizer::PoseOptimization(Frame* pFrame)
{
    const int N = pFrame->N;
    std::vector<int> index;
    std::vector<double> Xw, obs;
    std::vector<float> invSigma2;
    index.reserve(N); Xw.reserve((size_t)3 * N); obs.reserve((size_t)3 * N); invSigma2.reserve(N);
    {
        unique_lock<mutex> lock(MapPoint::mGlobalMutex);                         // :274
        for (int i = 0; i < N; i++)
        {
            MapPoint* pMP = pFrame->mvpMapPoints[i];
            if (!pMP)
                continue;
            pFrame->mvbOutlier[i] = false;                                       // :284, :319
            const cv::KeyPoint& kpUn = pFrame->mvKeysUn[i];
            const cv::Mat X = pMP->GetWorldPos();
            index.push_back(i);
            for (int k = 0; k < 3; k++) Xw.push_back(X.at<float>(k));
            obs.push_back(kpUn.pt.x); obs.push_back(kpUn.pt.y); obs.push_back(pFrame->mvuRight[i]);   // a negative uRight selects the monocular edge (:281)
            invSigma2.push_back(pFrame->mvInvLevelSigma2[kpUn.octave]);
        }
    }
    const int nInitialCorrespondences = (int)index.size();
    if (nInitialCorrespondences < 3)                                             // :355-356
        return 0;

    thread_local PoseHandle handle;
    orbx_pose* h = handle.fit(nInitialCorrespondences);
    if (!h)
        return 0;
    orbx_pose_problem P = orbx_pose_problem();
    P.n = nInitialCorrespondences; P.Xw = Xw.data(); P.obs = obs.data(); P.inv_sigma2 = invSigma2.data();
    poseOf(Converter::toSE3Quat(pFrame->mTcw), P.pose);                          // :254, :366
    P.fx = pFrame->fx; P.fy = pFrame->fy; P.cx = pFrame->cx; P.cy = pFrame->cy; P.bf = pFrame->mbf;
    std::vector<uint8_t> outlier(nInitialCorrespondences, 0);
    orbx_pose_result R = orbx_pose_result();
    R.outlier = outlier.data();
    // four rounds of optimize(10) with re-classification (:358-421)
    if (orbxFailed(orbx_pose_optimize_host(h, &P, 1, &R), "orbx_pose_optimize_host"))
        return 0;
    for (int k = 0; k < nInitialCorrespondences; k++)
        pFrame->mvbOutlier[index[k]] = outlier[k] != 0;
    pFrame->SetPose(Converter::toCvMat(se3Of(R.pose)));                          // :424-428
    pFrame->nBadPoseOpt = R.n_bad;
    return R.n_inliers;
}

} // namespace ORB_SLAM2
