// ORB_SLAM2::ORBmatcher on top of the orbx C ABI, second part: the vocabulary-node matchers SearchByBoW (both overloads),
// SearchForTriangulation, and SearchForInitialization.  include/ORBmatcher.h stays as it is; build like ORBmatcher_orbx.cc
// (these definitions removed from / guarded in the reference's src/ORBmatcher.cc).
#include "ORBmatcher.h"

#include "orbx_adapter.h"

#include <algorithm>

using namespace std;

namespace ORB_SLAM2
{

namespace
{

// a DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>) as ascending node ids + CSR lists
struct FeatCsr
{
    std::vector<uint32_t> id;
    std::vector<int32_t> start, feat;
    explicit FeatCsr(const DBoW2::FeatureVector& fv)
    {
        start.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it)
        {
            id.push_back((uint32_t)it->first);
            for (size_t k = 0; k < it->second.size(); k++) feat.push_back((int32_t)it->second[k]);
            start.push_back((int32_t)feat.size());
        }
    }
};

orbx_bow_set bowSet(int n, const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& desc, const std::vector<float>& uRight,
                    const std::vector<uint8_t>& valid, const FeatCsr& fv)
{
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint is the 28-byte record of orbx_keypoint");
    orbx_bow_set s = orbx_bow_set();
    s.n = n;
    s.keys_un = reinterpret_cast<const orbx_keypoint*>(keysUn.data());
    s.desc = desc.data;
    s.u_right = uRight.data();
    s.valid = valid.data();
    s.n_nodes = (int32_t)fv.id.size();
    s.node_id = fv.id.data(); s.node_start = fv.start.data(); s.node_feat = fv.feat.data();
    return s;
}

std::vector<uint8_t> goodMapPoints(const std::vector<MapPoint*>& v)
{
    std::vector<uint8_t> ok(v.size(), 0);
    for (size_t i = 0; i < v.size(); i++) ok[i] = v[i] && !v[i]->isBad();
    return ok;
}
} // namespace

// replaces ORBmatcher.cc:159-288
int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches)
{
    const vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    const FeatCsr fvKF(pKF->mFeatVec), fvF(F.mFeatVec);
    const std::vector<uint8_t> validKF = goodMapPoints(vpMapPointsKF), validF(F.N, 1);
    orbx_bucket_job job = orbx_bucket_job();
    job.a = bowSet(pKF->N, pKF->mvKeysUn, pKF->mDescriptors, pKF->mvuRight, validKF, fvKF);
    job.b = bowSet(F.N, F.mvKeysUn, F.mDescriptors, F.mvuRight, validF, fvF);
    job.mode = 0; job.nnratio = mfNNratio; job.check_ori = mbCheckOrientation;
    job.sigma2_b = F.mvLevelSigma2.data(); job.scale_b = F.mvScaleFactors.data(); job.nlevels = F.mnScaleLevels;
    std::vector<int32_t> matchKF(pKF->N > 0 ? pKF->N : 1, -1);
    int32_t nmatches = 0;
    orbx_matcher* m = orbxMatcherOfThisThread(std::max(job.a.n, job.b.n), std::max(job.a.n, job.b.n));
    if (!m || orbxFailed(orbx_match_buckets_host(m, &job, matchKF.data(), &nmatches), "SearchByBoW(KeyFrame, Frame)"))
        return 0;
    for (int i = 0; i < pKF->N; i++)
        if (matchKF[i] >= 0)
            vpMapPointMatches[matchKF[i]] = vpMapPointsKF[i];                    // :256
    return nmatches;
}

// replaces ORBmatcher.cc:522-655
int ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12)
{
    const vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    const FeatCsr fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    const std::vector<uint8_t> valid1 = goodMapPoints(vpMapPoints1), valid2 = goodMapPoints(vpMapPoints2);
    orbx_bucket_job job = orbx_bucket_job();
    job.a = bowSet(pKF1->N, pKF1->mvKeysUn, pKF1->mDescriptors, pKF1->mvuRight, valid1, fv1);
    job.b = bowSet(pKF2->N, pKF2->mvKeysUn, pKF2->mDescriptors, pKF2->mvuRight, valid2, fv2);
    job.mode = 1; job.nnratio = mfNNratio; job.check_ori = mbCheckOrientation;
    job.sigma2_b = pKF2->mvLevelSigma2.data(); job.scale_b = pKF2->mvScaleFactors.data(); job.nlevels = pKF2->mnScaleLevels;
    std::vector<int32_t> match12(pKF1->N > 0 ? pKF1->N : 1, -1);
    int32_t nmatches = 0;
    orbx_matcher* m = orbxMatcherOfThisThread(std::max(job.a.n, job.b.n), std::max(job.a.n, job.b.n));
    if (!m || orbxFailed(orbx_match_buckets_host(m, &job, match12.data(), &nmatches), "SearchByBoW(KeyFrame, KeyFrame)"))
        return 0;
    for (int i = 0; i < pKF1->N; i++)
        if (match12[i] >= 0)
            vpMatches12[i] = vpMapPoints2[match12[i]];                           // :610
    return nmatches;
}

// replaces ORBmatcher.cc:657-823
int ORBmatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, vector<pair<size_t, size_t> >& vMatchedPairs,
                                       const bool bOnlyStereo)
{
    // epipole in the second image, :663-670; C2 = R2w*Cw+t2w as sequential float multiply-adds (what the cv::Mat product does)
    const cv::Mat Cw = pKF1->GetCameraCenter(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    float C2[3];
    for (int r = 0; r < 3; r++)
        C2[r] = ((R2w.at<float>(r, 0) * Cw.at<float>(0) + R2w.at<float>(r, 1) * Cw.at<float>(1)) + R2w.at<float>(r, 2) * Cw.at<float>(2)) +
                t2w.at<float>(r);
    const float invz = 1.0f / C2[2];
    const float ex = pKF2->fx * C2[0] * invz + pKF2->cx;
    const float ey = pKF2->fy * C2[1] * invz + pKF2->cy;

    // only keypoints WITHOUT a map point take part (:699-702, :727-730)
    std::vector<uint8_t> free1(pKF1->N, 0), free2(pKF2->N, 0);
    for (int i = 0; i < pKF1->N; i++) free1[i] = pKF1->GetMapPoint(i) == NULL;
    for (int i = 0; i < pKF2->N; i++) free2[i] = pKF2->GetMapPoint(i) == NULL;
    const FeatCsr fv1(pKF1->mFeatVec), fv2(pKF2->mFeatVec);
    orbx_bucket_job job = orbx_bucket_job();
    job.a = bowSet(pKF1->N, pKF1->mvKeysUn, pKF1->mDescriptors, pKF1->mvuRight, free1, fv1);
    job.b = bowSet(pKF2->N, pKF2->mvKeysUn, pKF2->mDescriptors, pKF2->mvuRight, free2, fv2);
    job.mode = 2; job.nnratio = mfNNratio; job.check_ori = mbCheckOrientation; job.only_stereo = bOnlyStereo;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) job.F12[3 * r + c] = F12.at<float>(r, c);
    job.ex = ex; job.ey = ey;
    job.sigma2_b = pKF2->mvLevelSigma2.data(); job.scale_b = pKF2->mvScaleFactors.data(); job.nlevels = pKF2->mnScaleLevels;
    std::vector<int32_t> match12(pKF1->N > 0 ? pKF1->N : 1, -1);
    int32_t nmatches = 0;
    vMatchedPairs.clear();
    orbx_matcher* m = orbxMatcherOfThisThread(std::max(job.a.n, job.b.n), std::max(job.a.n, job.b.n));
    if (!m || orbxFailed(orbx_match_buckets_host(m, &job, match12.data(), &nmatches), "SearchForTriangulation"))
        return 0;
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < pKF1->N; i++)                                            // :812-820
        if (match12[i] >= 0)
            vMatchedPairs.push_back(make_pair((size_t)i, (size_t)match12[i]));
    return nmatches;
}

// replaces ORBmatcher.cc:405-520 (monocular initialisation)
int ORBmatcher::SearchForInitialization(Frame& F1, Frame& F2, vector<cv::Point2f>& vbPrevMatched, vector<int>& vnMatches12, int windowSize)
{
    vnMatches12 = vector<int>(F1.mvKeysUn.size(), -1);
    std::vector<uint8_t> none1(F1.N, 0), none2(F2.N, 0);
    orbx_frame_view v[2];
    Frame* fr[2] = {&F1, &F2};
    for (int k = 0; k < 2; k++)
    {
        const Frame& F = *fr[k];
        v[k] = orbx_frame_view();
        v[k].n = F.N;
        v[k].keys_un = reinterpret_cast<const orbx_keypoint*>(F.mvKeysUn.data());
        v[k].desc = F.mDescriptors.data;
        v[k].u_right = F.mvuRight.data();
        v[k].claimed = k == 0 ? none1.data() : none2.data();
        v[k].min_x = Frame::mnMinX; v[k].min_y = Frame::mnMinY; v[k].max_x = Frame::mnMaxX; v[k].max_y = Frame::mnMaxY;
        v[k].grid_w_inv = Frame::mfGridElementWidthInv; v[k].grid_h_inv = Frame::mfGridElementHeightInv;
        v[k].fx = Frame::fx; v[k].fy = Frame::fy; v[k].cx = Frame::cx; v[k].cy = Frame::cy; v[k].bf = F.mbf; v[k].b = F.mb;
        v[k].scale_factors = F.mvScaleFactors.data();
        v[k].nlevels = F.mnScaleLevels;
    }
    static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "vbPrevMatched is passed as (x, y) float pairs");
    std::vector<int32_t> match12(F1.N > 0 ? F1.N : 1, -1);
    int32_t nmatches = 0;
    orbx_matcher* m = orbxMatcherOfThisThread(std::max(F1.N, F2.N), std::max(F1.N, F2.N));
    if (!m || orbxFailed(orbx_match_initialization_host(m, &v[0], &v[1], reinterpret_cast<const float*>(vbPrevMatched.data()), windowSize,
                                                        mfNNratio, mbCheckOrientation, match12.data(), &nmatches), "SearchForInitialization"))
        nmatches = 0;                                                        // match12 stays all -1
    for (int i1 = 0; i1 < F1.N; i1++)
    {
        vnMatches12[i1] = match12[i1];
        if (match12[i1] >= 0)
            vbPrevMatched[i1] = F2.mvKeysUn[match12[i1]].pt;                     // :513-516
    }
    return nmatches;
}

} // namespace ORB_SLAM2
