// Shared by the adapter files (the reference's C++ classes on top of the orbx C ABI).
//
// Error behaviour: the reference has none -- void / count returns, no exceptions, nothing in Tracking.cc / LocalMapping.cc /
// LoopClosing.cc catches anything.  A failure of the device library (no device, out of memory, a CUDA error) is therefore reported
// on stderr and the member returns what the reference returns when it finds nothing (0 matches, empty keypoints, the pose left
// as it was); nothing is thrown into the callers.
//
// Device: ORBX_DEVICE in the environment selects the CUDA device of every handle the adapters create (default 0).
#ifndef ORBX_ADAPTER_H
#define ORBX_ADAPTER_H

#include <orbx.h>

#include <cstdio>
#include <cstdlib>

namespace ORB_SLAM2
{

inline int orbxDevice()
{
    static const int device = std::getenv("ORBX_DEVICE") ? std::atoi(std::getenv("ORBX_DEVICE")) : 0;
    return device;
}

// true if the call failed (and says so on stderr)
inline bool orbxFailed(orbx_status s, const char* where)
{
    if (s == ORBX_OK)
        return false;
    std::fprintf(stderr, "orbx: %s failed (%d): %s\n", where, (int)s, orbx_last_error());
    return true;
}

// ORBmatcher objects are stack temporaries in the reference (one per call); the device scratch lives per calling thread and is
// shared by the three ORBmatcher adapter files.  It is sized by the largest call seen so far: Tracking passes all of
// mvpLocalMapPoints (80+ local keyframes hold more than 8192 points), Fuse / loop closing pass their candidate lists.  NULL if the
// device library cannot provide it.
orbx_matcher* orbxMatcherOfThisThread(int nKeypoints, int nPoints);

} // namespace ORB_SLAM2

#endif
