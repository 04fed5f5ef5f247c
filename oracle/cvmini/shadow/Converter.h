// Shadow of the reference's include/Converter.h for the oracle/_ref build of src/Frame.cc: the real header pulls in Eigen and
// g2o, which are absent here.  Only the one function that Frame::ComputeBoW uses is declared (defined in orbmatcher_ref_shim.cpp:
// the rows of the descriptor matrix as a vector, like Converter.cc:31-39).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <opencv2/core/core.hpp>
#include <vector>
namespace ORB_SLAM2 {
class Converter {
public:
    static std::vector<cv::Mat> toDescriptorVector(const cv::Mat &Descriptors);
};
}
