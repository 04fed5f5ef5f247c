// Shadow in front of the reference's include/Optimizer.h for the oracle/_ref builds (TEST INFRASTRUCTURE ONLY).
// include/Optimizer.h:28 includes LoopClosing.h only for the two container typedefs in OptimizeEssentialGraph's signature
// (include/LoopClosing.h:47-49); the real LoopClosing.h drags in Tracking.h -> Viewer.h / Planning.h / OctomapBuilder.h (Pangolin,
// OMPL, octomap), none of which exist in this image.  This file declares those two typedefs exactly as the reference does, marks
// LoopClosing.h as already included, and then hands over to the reference's own, unmodified Optimizer.h.
#ifndef ORBX_SHADOW_OPTIMIZER_H
#define ORBX_SHADOW_OPTIMIZER_H
#ifndef LOOPCLOSING_H
#define LOOPCLOSING_H
#include "KeyFrame.h"
#include "Map.h"
#include <map>
#include <set>
#include <Eigen/StdVector>
#include "Thirdparty/g2o/g2o/types/types_seven_dof_expmap.h"
namespace ORB_SLAM2 {
class LoopClosing {
public:
    typedef std::pair<std::set<KeyFrame*>, int> ConsistentGroup;
    typedef std::map<KeyFrame*, g2o::Sim3, std::less<KeyFrame*>, Eigen::aligned_allocator<std::pair<const KeyFrame*, g2o::Sim3> > > KeyFrameAndPose;
};
}
#endif
#include_next "Optimizer.h"
#endif
