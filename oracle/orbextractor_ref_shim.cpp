// C entry points around the REFERENCE's own ORB_SLAM2::ORBextractor (src/ORBextractor.cc, compiled from where it lies
// under /root/reference against the OpenCV stand-in oracle/cvmini; see that header for what is real and what is stood
// in).  TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref_extractor.py and tools/make_ref_golden.py drive it to pin the
// oracle (oracle/extract_oracle.c) against the reference's code.
//
// Heap order: DistributeOctTree sorts (size, ExtractorNode*) pairs (ORBextractor.cc:684), so among nodes with equally
// many keypoints the expansion order — and through it a few of the returned keypoints — depends on where malloc put
// the list nodes.  The oracle and the CUDA path define that order as CREATION order.  To compare like with like this
// library (and only it: linked with -Bsymbolic) replaces operator new with a bump allocator over one reserved region,
// so that inside the reference's code a later allocation always has the larger address.  Single-threaded by design.
#include "ORBextractor.h"
#include <cstring>
#include "ref_bump_alloc.h"

extern "C" {
void *orbref_extractor_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th) {
    orbref_arena_retain();
    return new ORB_SLAM2::ORBextractor(nfeatures, scale_factor, nlevels, ini_th, min_th);
}
void orbref_extractor_destroy(void *h) {
    if (!h) return;
    delete static_cast<ORB_SLAM2::ORBextractor *>(h);
    orbref_arena_release();
}

// ORBextractor::operator(): returns the number of keypoints; the first min(n, cap) are copied out (28-byte cv::KeyPoint
// records, 32-byte descriptor rows)
int orbref_extract(void *h, const unsigned char *img, int w, int height, int stride, void *kps, unsigned char *desc, int cap) {
    ORB_SLAM2::ORBextractor &e = *static_cast<ORB_SLAM2::ORBextractor *>(h);
    cv::Mat image = (w > 0 && height > 0) ? cv::Mat(height, w, CV_8UC1, const_cast<unsigned char *>(img), (size_t)stride) : cv::Mat();
    std::vector<cv::KeyPoint> keys;
    cv::Mat d;
    e(image, cv::Mat(), keys, d);
    const int n = (int)keys.size(), m = n < cap ? n : cap;
    static_assert(sizeof(cv::KeyPoint) == 28, "cv::KeyPoint layout");
    if (m > 0) {
        std::memcpy(kps, keys.data(), (size_t)m * sizeof(cv::KeyPoint));
        for (int i = 0; i < m; i++) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    }
    return n;
}
// getters (include/ORBextractor.h:65-85) and the pyramid (mvImagePyramid) of the last call
int orbref_levels(void *h) { return static_cast<ORB_SLAM2::ORBextractor *>(h)->GetLevels(); }
void orbref_tables(void *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2) {
    ORB_SLAM2::ORBextractor &e = *static_cast<ORB_SLAM2::ORBextractor *>(h);
    const std::vector<float> a = e.GetScaleFactors(), b = e.GetInverseScaleFactors(), c = e.GetScaleSigmaSquares(),
                             d = e.GetInverseScaleSigmaSquares();
    for (size_t i = 0; i < a.size(); i++) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}
int orbref_level(void *h, int level, int *w, int *height, int *stride, const unsigned char **ptr) {
    ORB_SLAM2::ORBextractor &e = *static_cast<ORB_SLAM2::ORBextractor *>(h);
    if (level < 0 || level >= (int)e.mvImagePyramid.size() || e.mvImagePyramid[level].empty()) return -1;
    const cv::Mat &m = e.mvImagePyramid[level];
    *w = m.cols; *height = m.rows; *stride = (int)m.step; *ptr = m.data;
    return 0;
}
}
