// ORB_SLAM2::ORBmatcher on top of the orbx C ABI, third part: the keyframe projection searches
//   SearchByProjection(CurrentFrame, KeyFrame*, sAlreadyFound, th, ORBdist)   relocalisation        ORBmatcher.cc:1472-1599
//   SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)               loop closing          :290-403
//   Fuse(KeyFrame*, vpMapPoints, th)                                          local mapping         :825-975
//   Fuse(KeyFrame*, Scw, vpPoints, th, vpReplacePoint)                        loop correction       :977-1100
//   SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th)                  loop closing          :1102-1326
// The per-point host geometry (projection, IsInImage, distance-invariance and viewing-angle gates, PredictScale) is evaluated
// here with the reference's own cv::Mat expressions and MapPoint / KeyFrame methods, so it is the reference's arithmetic by
// construction; the window search + Hamming distances of ALL points of a call are one orbx_match_window_host /
// orbx_match_projection_keyframe_host call; map mutations are then applied on the host in the reference's order.
// include/ORBmatcher.h stays as it is; build like ORBmatcher_orbx.cc.
#include "ORBmatcher.h"

#include "orbx_adapter.h"

#include <climits>
#include <cmath>

using namespace std;

namespace ORB_SLAM2
{

namespace
{

// a KeyFrame as the window searches read it (KeyFrame::GetFeaturesInArea has no level filter, KeyFrame.cc:630-669)
orbx_frame_view viewOf(const KeyFrame* pKF, const std::vector<uint8_t>& claimed)
{
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint is the 28-byte record of orbx_keypoint");
    orbx_frame_view v = orbx_frame_view();
    v.n = pKF->N;
    v.keys_un = reinterpret_cast<const orbx_keypoint*>(pKF->mvKeysUn.data());
    v.desc = pKF->mDescriptors.data;
    v.u_right = pKF->mvuRight.data();
    v.claimed = claimed.data();
    v.min_x = pKF->mnMinX; v.min_y = pKF->mnMinY; v.max_x = pKF->mnMaxX; v.max_y = pKF->mnMaxY;
    v.grid_w_inv = pKF->mfGridElementWidthInv; v.grid_h_inv = pKF->mfGridElementHeightInv;
    v.fx = pKF->fx; v.fy = pKF->fy; v.cx = pKF->cx; v.cy = pKF->cy; v.bf = pKF->mbf; v.b = pKF->mb;
    v.scale_factors = pKF->mvScaleFactors.data();
    v.nlevels = pKF->mnScaleLevels;
    return v;
}

// the candidates of one call: window points in the caller's order, their descriptors, and which caller index each one is
struct Candidates
{
    std::vector<orbx_window_point> pts;
    std::vector<uint8_t> desc;
    std::vector<int> source;
    void add(int src, float u, float v, float ur, float radius, int level, MapPoint* pMP)
    {
        orbx_window_point w = orbx_window_point();
        w.u = u; w.v = v; w.ur = ur; w.radius = radius;
        w.min_level = level - 1; w.max_level = level;
        w.valid = 1;
        pts.push_back(w);
        const cv::Mat d = pMP->GetDescriptor();
        desc.insert(desc.end(), d.data, d.data + 32);
        source.push_back(src);
    }
};

// one window search for all candidates; best[i] = keypoint index or -1 (no candidate or distance > maxDist)
std::vector<int32_t> search(const KeyFrame* pKF, const std::vector<uint8_t>& claimed, const Candidates& c, int flags, int maxDist)
{
    const int n = (int)c.pts.size();
    std::vector<int32_t> best(n > 0 ? n : 1, -1), dist(n > 0 ? n : 1, 256);
    if (n == 0)
        return best;
    const orbx_frame_view view = viewOf(pKF, claimed);
    int32_t accepted = 0;
    orbx_matcher* m = orbxMatcherOfThisThread(pKF->N, n);
    if (!m || orbxFailed(orbx_match_window_host(m, &view, n, c.pts.data(), c.desc.data(), flags, pKF->mvInvLevelSigma2.data(), maxDist, best.data(),
                                                dist.data(), &accepted), "window search (Fuse / SearchBySim3 / SearchByProjection(KeyFrame, Scw))"))
        best.assign(best.size(), -1);                                        // nothing found: the callers then change nothing
    return best;
}

// projection + gates shared by SearchByProjection(KF, Scw, ...), Fuse(KF, ...) and Fuse(KF, Scw, ...): returns false if the point
// is discarded before the window search
bool projectIntoKeyFrame(KeyFrame* pKF, MapPoint* pMP, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, float th, float& u,
                         float& v, float& ur, float& radius, int& level)
{
    const cv::Mat p3Dw = pMP->GetWorldPos();
    const cv::Mat p3Dc = Rcw * p3Dw + tcw;
    if (p3Dc.at<float>(2) < 0.0f)                                    // depth must be positive
        return false;
    const float invz = 1 / p3Dc.at<float>(2);
    const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
    u = pKF->fx * x + pKF->cx;
    v = pKF->fy * y + pKF->cy;
    if (!pKF->IsInImage(u, v))
        return false;
    ur = u - pKF->mbf * invz;
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    const cv::Mat PO = p3Dw - Ow;
    const float dist3D = cv::norm(PO);
    if (dist3D < minDistance || dist3D > maxDistance)
        return false;
    const cv::Mat Pn = pMP->GetNormal();
    if (PO.dot(Pn) < 0.5 * dist3D)                                   // viewing angle below 60 degrees
        return false;
    level = pMP->PredictScale(dist3D, pKF);
    radius = th * pKF->mvScaleFactors[level];
    return true;
}

void decomposeSim3(const cv::Mat& Scw, cv::Mat& Rcw, cv::Mat& tcw, cv::Mat& Ow)
{
    const cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    Rcw = sRcw / scw;
    tcw = Scw.rowRange(0, 3).col(3) / scw;
    Ow = -Rcw.t() * tcw;
}
} // namespace

// replaces ORBmatcher.cc:290-403
int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, vector<MapPoint*>& vpMatched, int th)
{
    cv::Mat Rcw, tcw, Ow;
    decomposeSim3(Scw, Rcw, tcw, Ow);
    set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    Candidates c;
    for (int iMP = 0, iendMP = (int)vpPoints.size(); iMP < iendMP; iMP++)
    {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP))
            continue;
        float u, v, ur, radius; int level;
        if (projectIntoKeyFrame(pKF, pMP, Rcw, tcw, Ow, (float)th, u, v, ur, radius, level))
            c.add(iMP, u, v, ur, radius, level, pMP);
    }
    // keypoints that already carry a match are skipped, and so are those an earlier point of this call took (:377-378, :396): flags = 2
    std::vector<uint8_t> claimed(pKF->N, 0);
    for (int k = 0; k < pKF->N && k < (int)vpMatched.size(); k++) claimed[k] = vpMatched[k] != NULL;
    const std::vector<int32_t> best = search(pKF, claimed, c, 2, TH_LOW);
    int nmatches = 0;
    for (size_t i = 0; i < c.source.size(); i++)
        if (best[i] >= 0)
        {
            vpMatched[best[i]] = vpPoints[c.source[i]];
            nmatches++;
        }
    return nmatches;
}

// replaces ORBmatcher.cc:825-975
int ORBmatcher::Fuse(KeyFrame* pKF, const vector<MapPoint*>& vpMapPoints, const float th)
{
    const cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
    Candidates c;
    for (int i = 0, n = (int)vpMapPoints.size(); i < n; i++)
    {
        MapPoint* pMP = vpMapPoints[i];
        if (!pMP || pMP->isBad() || pMP->IsInKeyFrame(pKF))
            continue;
        float u, v, ur, radius; int level;
        if (projectIntoKeyFrame(pKF, pMP, Rcw, tcw, Ow, th, u, v, ur, radius, level))
            c.add(i, u, v, ur, radius, level, pMP);
    }
    // flags = 1: the reprojection chi2 gate of :905-930 inside the window scan; no claims
    const std::vector<uint8_t> none(pKF->N, 0);
    const std::vector<int32_t> best = search(pKF, none, c, 1, TH_LOW);
    int nFused = 0;
    for (size_t i = 0; i < c.source.size(); i++)
    {
        MapPoint* pMP = vpMapPoints[c.source[i]];
        // an earlier replacement of this call may have made the point bad or put it into the keyframe: the reference tests
        // that when the point's turn comes (:847-848), and the search result itself does not depend on the other points
        if (best[i] < 0 || pMP->isBad() || pMP->IsInKeyFrame(pKF))
            continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
        if (pMPinKF)
        {
            if (!pMPinKF->isBad())
            {
                if (pMPinKF->Observations() > pMP->Observations())
                    pMP->Replace(pMPinKF);
                else
                    pMPinKF->Replace(pMP);
            }
        }
        else
        {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// replaces ORBmatcher.cc:977-1100
int ORBmatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*>& vpPoints, float th, vector<MapPoint*>& vpReplacePoint)
{
    cv::Mat Rcw, tcw, Ow;
    decomposeSim3(Scw, Rcw, tcw, Ow);
    const set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
    Candidates c;
    for (int iMP = 0, n = (int)vpPoints.size(); iMP < n; iMP++)
    {
        MapPoint* pMP = vpPoints[iMP];
        if (pMP->isBad() || spAlreadyFound.count(pMP))
            continue;
        float u, v, ur, radius; int level;
        if (projectIntoKeyFrame(pKF, pMP, Rcw, tcw, Ow, th, u, v, ur, radius, level))
            c.add(iMP, u, v, ur, radius, level, pMP);
    }
    const std::vector<uint8_t> none(pKF->N, 0);
    const std::vector<int32_t> best = search(pKF, none, c, 0, TH_LOW);
    int nFused = 0;
    for (size_t i = 0; i < c.source.size(); i++)
    {
        if (best[i] < 0)
            continue;
        MapPoint* pMP = vpPoints[c.source[i]];
        MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
        if (pMPinKF)
        {
            if (!pMPinKF->isBad())
                vpReplacePoint[c.source[i]] = pMPinKF;
        }
        else
        {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

// replaces ORBmatcher.cc:1102-1326
int ORBmatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th)
{
    const cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    const cv::Mat sR12 = s12 * R12;
    const cv::Mat sR21 = (1.0 / s12) * R12.t();
    const cv::Mat t21 = -sR21 * t12;
    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++)
    {
        MapPoint* pMP = vpMatches12[i];
        if (pMP)
        {
            vbAlreadyMatched1[i] = true;
            const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
            if (idx2 >= 0 && idx2 < N2)
                vbAlreadyMatched2[idx2] = true;
        }
    }
    // one direction: the not yet matched map points of one keyframe, taken to its camera with (Rw, tw), to the other camera with
    // (sR, t), projected with pKF1's intrinsics (the reference reads fx, fy, cx, cy from pKF1 for both directions, :1106-1109) and
    // searched in `to`
    auto direction = [&](KeyFrame* to, const vector<MapPoint*>& pts, const vector<bool>& already, const cv::Mat& Rw, const cv::Mat& tw,
                         const cv::Mat& sR, const cv::Mat& t) {
        Candidates c;
        for (int i = 0, n = (int)pts.size(); i < n; i++)
        {
            MapPoint* pMP = pts[i];
            if (!pMP || already[i] || pMP->isBad())
                continue;
            const cv::Mat p3Dw = pMP->GetWorldPos();
            const cv::Mat p3Dfrom = Rw * p3Dw + tw;
            const cv::Mat p3Dto = sR * p3Dfrom + t;
            if (p3Dto.at<float>(2) < 0.0)
                continue;
            const float invz = 1.0 / p3Dto.at<float>(2);
            const float x = p3Dto.at<float>(0) * invz, y = p3Dto.at<float>(1) * invz;
            const float u = pKF1->fx * x + pKF1->cx, v = pKF1->fy * y + pKF1->cy;
            if (!to->IsInImage(u, v))
                continue;
            const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
            const float dist3D = cv::norm(p3Dto);
            if (dist3D < minDistance || dist3D > maxDistance)
                continue;
            const int level = pMP->PredictScale(dist3D, to);
            c.add(i, u, v, -1.0f, th * to->mvScaleFactors[level], level, pMP);
        }
        const std::vector<uint8_t> none(to->N, 0);
        const std::vector<int32_t> best = search(to, none, c, 0, TH_HIGH);
        std::vector<int> match(pts.size(), -1);
        for (size_t k = 0; k < c.source.size(); k++) match[c.source[k]] = best[k];
        return match;
    };
    const std::vector<int> vnMatch1 = direction(pKF2, vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21);
    const std::vector<int> vnMatch2 = direction(pKF1, vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12);
    int nFound = 0;                                                  // mutual agreement, :1300-1323
    for (int i1 = 0; i1 < N1; i1++)
    {
        const int idx2 = vnMatch1[i1];
        if (idx2 >= 0 && vnMatch2[idx2] == i1)
        {
            vpMatches12[i1] = vpMapPoints2[idx2];
            nFound++;
        }
    }
    return nFound;
}

// replaces ORBmatcher.cc:1472-1599 (relocalisation)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist)
{
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    const vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int n = (int)vpMPs.size();
    // the device projects again with the same float operations (:1498-1506) and tests the image bounds; the distance gate and
    // PredictScale (:1513-1523) are evaluated here and passed as valid / octave
    std::vector<orbx_last_point> pts(n > 0 ? n : 1);
    std::vector<uint8_t> desc((size_t)32 * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++)
    {
        MapPoint* pMP = vpMPs[i];
        orbx_last_point& p = pts[i];
        p = orbx_last_point();
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP))
            continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        const cv::Mat PO = x3Dw - Ow;
        const float dist3D = cv::norm(PO);
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        if (dist3D < minDistance || dist3D > maxDistance)
            continue;
        p.x = x3Dw.at<float>(0); p.y = x3Dw.at<float>(1); p.z = x3Dw.at<float>(2);
        p.angle = pKF->mvKeysUn[i].angle;
        p.octave = pMP->PredictScale(dist3D, &CurrentFrame);
        p.valid = 1;
        p.blocks = 1;
        const cv::Mat d = pMP->GetDescriptor();
        for (int k = 0; k < 32; k++) desc[(size_t)32 * i + k] = d.data[k];
    }
    // this overload skips every keypoint that already has a map point, whatever its observations (:1532-1533)
    std::vector<uint8_t> claimed(CurrentFrame.N, 0);
    for (int k = 0; k < CurrentFrame.N; k++) claimed[k] = CurrentFrame.mvpMapPoints[k] != NULL;
    orbx_frame_view view = orbx_frame_view();
    view.n = CurrentFrame.N;
    view.keys_un = reinterpret_cast<const orbx_keypoint*>(CurrentFrame.mvKeysUn.data());
    view.desc = CurrentFrame.mDescriptors.data;
    view.u_right = CurrentFrame.mvuRight.data();
    view.claimed = claimed.data();
    view.min_x = Frame::mnMinX; view.min_y = Frame::mnMinY; view.max_x = Frame::mnMaxX; view.max_y = Frame::mnMaxY;
    view.grid_w_inv = Frame::mfGridElementWidthInv; view.grid_h_inv = Frame::mfGridElementHeightInv;
    view.fx = Frame::fx; view.fy = Frame::fy; view.cx = Frame::cx; view.cy = Frame::cy; view.bf = CurrentFrame.mbf; view.b = CurrentFrame.mb;
    view.scale_factors = CurrentFrame.mvScaleFactors.data();
    view.nlevels = CurrentFrame.mnScaleLevels;
    float R[9], t[3];
    for (int r = 0; r < 3; r++)
    {
        for (int cc = 0; cc < 3; cc++) R[3 * r + cc] = Rcw.at<float>(r, cc);
        t[r] = tcw.at<float>(r);
    }
    std::vector<int32_t> match(CurrentFrame.N > 0 ? CurrentFrame.N : 1, -1);
    int32_t nmatches = 0;
    orbx_matcher* m = orbxMatcherOfThisThread(CurrentFrame.N, n);
    if (!m || orbxFailed(orbx_match_projection_keyframe_host(m, &view, n, pts.data(), desc.data(), R, t, th, ORBdist, mbCheckOrientation, match.data(),
                                                             &nmatches), "SearchByProjection(Frame, KeyFrame)"))
        return 0;
    for (int k = 0; k < CurrentFrame.N; k++)
        if (match[k] >= 0)
            CurrentFrame.mvpMapPoints[k] = vpMPs[match[k]];
    return nmatches;
}

} // namespace ORB_SLAM2
