// ORB_SLAM2::ORBmatcher on top of the orbx C ABI (include/orbx.h): the constructor, DescriptorDistance and the two
// SearchByProjection overloads Tracking calls on every frame (TrackWithMotionModel, Tracking.cc:857-880; SearchLocalPoints,
// Tracking.cc:1105-1116).  include/ORBmatcher.h stays as it is.
//
// Build: compile this file next to the reference's src/ORBmatcher.cc with the definitions of these four members removed
// there (or guarded by #ifndef ORBX_ADAPTER); the other overloads follow the same gather -> C call -> write-back pattern
// (table in INTEGRATION.md) and keep using the reference's code until they are moved over.
#include "ORBmatcher.h"

#include "orbx_adapter.h"

#include <algorithm>
#include <cmath>

using namespace std;

namespace ORB_SLAM2
{

const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

orbx_matcher* orbxMatcherOfThisThread(int nKeypoints, int nPoints)
{
    struct Scratch
    {
        orbx_matcher* m = nullptr;
        int kp = 0, pts = 0;
        ~Scratch() { if (m) orbx_matcher_destroy(m); }
    };
    thread_local Scratch sc;
    if (sc.m && nKeypoints <= sc.kp && nPoints <= sc.pts)
        return sc.m;
    if (sc.m) { orbx_matcher_destroy(sc.m); sc.m = nullptr; }
    sc.kp = std::max(8192, nKeypoints + nKeypoints / 2);
    sc.pts = std::max(8192, nPoints + nPoints / 2);
    if (orbxFailed(orbx_matcher_create(&sc.m, sc.kp, sc.pts, 1, orbxDevice()), "orbx_matcher_create"))
    {
        sc.m = nullptr;
        sc.kp = sc.pts = 0;
    }
    return sc.m;
}

namespace
{
// the Frame members the matchers read, as the POD view of include/orbx.h; `claimed` must outlive the call
orbx_frame_view viewOf(const Frame& F, std::vector<uint8_t>& claimed, bool byObservations)
{
    claimed.assign(F.N, 0);
    for (int i = 0; i < F.N; i++)
    {
        MapPoint* p = F.mvpMapPoints[i];
        claimed[i] = p && (!byObservations || p->Observations() > 0);        // ORBmatcher.cc:87-89, :1395-1397
    }
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint is the 28-byte record of orbx_keypoint");
    orbx_frame_view v = orbx_frame_view();
    v.n = F.N;
    v.keys_un = reinterpret_cast<const orbx_keypoint*>(F.mvKeysUn.data());
    v.desc = F.mDescriptors.data;
    v.u_right = F.mvuRight.data();
    v.claimed = claimed.data();
    v.min_x = Frame::mnMinX; v.min_y = Frame::mnMinY; v.max_x = Frame::mnMaxX; v.max_y = Frame::mnMaxY;
    v.grid_w_inv = Frame::mfGridElementWidthInv; v.grid_h_inv = Frame::mfGridElementHeightInv;
    v.fx = Frame::fx; v.fy = Frame::fy; v.cx = Frame::cx; v.cy = Frame::cy; v.bf = F.mbf; v.b = F.mb;
    v.scale_factors = F.mvScaleFactors.data();
    v.nlevels = F.mnScaleLevels;
    return v;
}
} // namespace

ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

// replaces ORBmatcher.cc:1647-1663
int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
{
    return orbx_hamming256(a.data, b.data);
}

// replaces ORBmatcher.cc:45-129 (track the local map): the points were prepared by Frame::isInFrustum
int ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)
{
    // only the points Frame::isInFrustum left in view take part (:53-57): they are packed, with an index back into vpMapPoints
    const int nAll = (int)vpMapPoints.size();
    std::vector<int> source;
    source.reserve(nAll);
    for (int i = 0; i < nAll; i++)
        if (vpMapPoints[i]->mbTrackInView && !vpMapPoints[i]->isBad())
            source.push_back(i);
    const int n = (int)source.size();
    std::vector<orbx_track_point> pts(n > 0 ? n : 1);
    std::vector<uint8_t> desc((size_t)32 * (n ? n : 1));
    for (int j = 0; j < n; j++)
    {
        MapPoint* pMP = vpMapPoints[source[j]];
        orbx_track_point& t = pts[j];
        t = orbx_track_point();
        t.proj_x = pMP->mTrackProjX; t.proj_y = pMP->mTrackProjY; t.proj_xr = pMP->mTrackProjXR;
        t.view_cos = pMP->mTrackViewCos;
        t.level = pMP->mnTrackScaleLevel;
        t.in_view = 1;
        t.blocks = pMP->Observations() > 0;
        const cv::Mat d = pMP->GetDescriptor();
        for (int k = 0; k < 32; k++) desc[(size_t)32 * j + k] = d.data[k];
    }
    std::vector<uint8_t> claimed;
    const orbx_frame_view view = viewOf(F, claimed, true);
    std::vector<int32_t> match(F.N > 0 ? F.N : 1, -1);
    int32_t nmatches = 0;
    orbx_matcher* m = orbxMatcherOfThisThread(F.N, n);
    // th is multiplied by RadiusByViewingCos and the level's scale factor inside (:63-69); the caller's bFactor logic
    // (`if(bFactor) r*=th`) is th itself
    if (!m || orbxFailed(orbx_match_projection_points_host(m, &view, n, pts.data(), desc.data(), th, mfNNratio, match.data(), &nmatches),
                         "SearchByProjection(Frame, MapPoints)"))
        return 0;
    for (int k = 0; k < F.N; k++)
        if (match[k] >= 0)
            F.mvpMapPoints[k] = vpMapPoints[source[match[k]]];                   // :123
    return nmatches;
}

// replaces ORBmatcher.cc:1328-1470 (track from the previous frame)
int ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
{
    // poses, :1338-1351.  cv::Mat CV_32F products are sequential float multiply-adds ((r0*x0 + r1*x1) + r2*x2) + t, written out
    float Rcw[9], tcw[3], twc[3], tlc[3];
    for (int r = 0; r < 3; r++)
    {
        for (int c = 0; c < 3; c++) Rcw[3 * r + c] = CurrentFrame.mTcw.at<float>(r, c);
        tcw[r] = CurrentFrame.mTcw.at<float>(r, 3);
    }
    for (int r = 0; r < 3; r++)                                                   // twc = -Rcw.t()*tcw
        twc[r] = -((Rcw[r] * tcw[0] + Rcw[3 + r] * tcw[1]) + Rcw[6 + r] * tcw[2]);
    for (int r = 0; r < 3; r++)                                                   // tlc = Rlw*twc+tlw
        tlc[r] = ((LastFrame.mTcw.at<float>(r, 0) * twc[0] + LastFrame.mTcw.at<float>(r, 1) * twc[1]) +
                  LastFrame.mTcw.at<float>(r, 2) * twc[2]) + LastFrame.mTcw.at<float>(r, 3);
    const bool bForward = tlc[2] > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc[2] > CurrentFrame.mb && !bMono;

    const int n = LastFrame.N;
    std::vector<orbx_last_point> pts(n);
    std::vector<uint8_t> desc((size_t)32 * (n ? n : 1));
    for (int i = 0; i < n; i++)
    {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        orbx_last_point& p = pts[i];
        p = orbx_last_point();
        if (!pMP || LastFrame.mvbOutlier[i])                                      // :1355-1359
            continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        p.x = x3Dw.at<float>(0); p.y = x3Dw.at<float>(1); p.z = x3Dw.at<float>(2);
        p.angle = LastFrame.mvKeysUn[i].angle;
        p.octave = LastFrame.mvKeys[i].octave;
        p.valid = 1;
        p.blocks = pMP->Observations() > 0;
        const cv::Mat d = pMP->GetDescriptor();
        for (int k = 0; k < 32; k++) desc[(size_t)32 * i + k] = d.data[k];
    }
    std::vector<uint8_t> claimed;
    const orbx_frame_view view = viewOf(CurrentFrame, claimed, true);
    std::vector<int32_t> match(CurrentFrame.N > 0 ? CurrentFrame.N : 1, -1);
    int32_t nmatches = 0;
    orbx_matcher* m = orbxMatcherOfThisThread(CurrentFrame.N, n);
    if (!m || orbxFailed(orbx_match_projection_frame_host(m, &view, n, pts.data(), desc.data(), Rcw, tcw, bForward, bBackward, th,
                                                          mbCheckOrientation, match.data(), &nmatches), "SearchByProjection(Cur, Last)"))
        return 0;
    // write-back: entries the rotation check rejected were set by this very call and come back as -1 (:1456-1462).
    // Precondition (holds at every call site of the reference: Tracking.cc:973 and :986 fill mvpMapPoints with NULL right before the
    // calls at :981 and :987): on entry every CurrentFrame.mvpMapPoints[k] is NULL or has Observations() > 0.  A keypoint that held a map
    // point WITHOUT observations, got matched and was then rejected would end as NULL in the reference and keeps its old pointer here.
    for (int k = 0; k < CurrentFrame.N; k++)
        if (match[k] >= 0)
            CurrentFrame.mvpMapPoints[k] = LastFrame.mvpMapPoints[match[k]];
    return nmatches;
}

} // namespace ORB_SLAM2
