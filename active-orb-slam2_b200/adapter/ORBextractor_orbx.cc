// ORB_SLAM2::ORBextractor on top of the orbx C ABI (include/orbx.h).
//
// Drop-in replacement for the reference's src/ORBextractor.cc: include/ORBextractor.h stays byte-for-byte as it is
// (reference include/ORBextractor.h:45-111), so Tracking.cc:120-126 and Frame.cc:37-43, 94-100, 276-282, 502, 592, 604,
// 609 compile and link unchanged.  Build: remove src/ORBextractor.cc from CMakeLists.txt, add this file, link -lorbx.
//
// The class has no spare member for the device handle and the header must not change, so handles live in a side table
// keyed by `this`.  The reference never destroys its extractors (Tracking.cc:120-126 allocates them once, ~ORBextractor
// is empty and inline in the header), so entries are never removed.
#include "ORBextractor.h"

#include "orbx_adapter.h"

#include <mutex>
#include <unordered_map>

namespace ORB_SLAM2
{

namespace
{
std::mutex gTableMutex;
std::unordered_map<const ORBextractor*, orbx_extractor*> gTable;

// device buffers are sized once per extractor; frames up to this size are accepted (KITTI 1241x376, EuRoC 752x480, TUM 640x480)
const int kMaxWidth = 1280, kMaxHeight = 1024;

orbx_extractor* handleOf(const ORBextractor* self)
{
    std::lock_guard<std::mutex> lock(gTableMutex);
    auto it = gTable.find(self);
    return it == gTable.end() ? nullptr : it->second;
}

} // namespace

// replaces ORBextractor::ORBextractor, src/ORBextractor.cc:410-470
ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST)
{
    orbx_extractor* h = nullptr;
    if (orbxFailed(orbx_extractor_create(&h, nfeatures, _scaleFactor, nlevels, iniThFAST, minThFAST, kMaxWidth, kMaxHeight, 1, orbxDevice()),
                   "orbx_extractor_create"))
        h = nullptr;                                      // operator() then returns no keypoints
    {
        std::lock_guard<std::mutex> lock(gTableMutex);
        gTable[this] = h;
    }
    mvScaleFactor.resize(nlevels);
    mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    if (!h || orbxFailed(orbx_extractor_tables(h, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                                               mnFeaturesPerLevel.data()), "orbx_extractor_tables"))
    {   // the scale tables of ORBextractor.cc:417-431, so that Frame's copies of them stay meaningful
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) { mvScaleFactor[i] = mvScaleFactor[i - 1] * scaleFactor; mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i]; }
        for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
    }
    mvImagePyramid.resize(nlevels);
    // `pattern` and `umax` (ORBextractor.cc:448-469) live in the device library; nothing on the host reads them
}

// replaces ORBextractor::operator(), src/ORBextractor.cc:1043-1105
void ORBextractor::operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                              cv::OutputArray _descriptors)
{
    if (_image.empty())                                   // :1046
        return;
    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1);                      // :1049

    orbx_extractor* h = handleOf(this);
    if (!h)
    {
        _keypoints.clear();
        _descriptors.release();
        return;
    }
    static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint is the 28-byte record of orbx_keypoint");
    const int cap = orbx_extractor_capacity(h);
    _keypoints.resize(cap);
    cv::Mat desc(cap, 32, CV_8U);
    const uint8_t* img = image.data;
    int32_t n = 0;
    if (orbxFailed(orbx_extractor_run_host(h, &img, 1, image.cols, image.rows, (int)image.step, reinterpret_cast<orbx_keypoint*>(_keypoints.data()),
                                           desc.data, &n), "ORBextractor::operator()"))
        n = 0;
    _keypoints.resize(n);
    if (n == 0)
        _descriptors.release();                           // :1064-1065
    else
        desc.rowRange(0, n).copyTo(_descriptors);         // :1068-1069

    // mvImagePyramid is public (ORBextractor.h:85).  Its only reader is Frame::ComputeStereoMatches (Frame.cc:502, 592, 609);
    // with adapter/Frame_stereo_orbx.cc that function reads the pyramids on the device and this copy can be compiled out.
#ifndef ORBX_ADAPTER_NO_HOST_PYRAMID
    for (int l = 0; l < nlevels; l++)
    {
        int w = 0, hgt = 0, pitch = 0;
        const uint8_t* d = nullptr;
        if (orbxFailed(orbx_extractor_pyramid(h, 0, l, &d, &w, &hgt, &pitch), "orbx_extractor_pyramid"))
            break;
        mvImagePyramid[l].create(hgt, w, CV_8U);
        if (orbxFailed(orbx_extractor_pyramid_host(h, 0, l, 0, mvImagePyramid[l].data, (int)mvImagePyramid[l].step), "orbx_extractor_pyramid_host"))
            break;
    }
#endif
}

// the device handle of an extractor, for the other adapter files (stereo association reads both pyramids in place)
orbx_extractor* orbxHandle(const ORBextractor* e) { return handleOf(e); }

// The protected helpers declared in the header (ComputePyramid, ComputeKeyPointsOctTree, DistributeOctTree,
// ComputeKeyPointsOld, ExtractorNode::DivideNode) are only called from inside src/ORBextractor.cc; nothing else in the
// tree references them, so they need no definition here.

} // namespace ORB_SLAM2
