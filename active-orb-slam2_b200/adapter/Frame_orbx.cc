// Frame::ComputeStereoMatches and Frame::ComputeBoW on top of the orbx C ABI (include/orbx.h); include/Frame.h stays as it is.
//
// Build: compile next to the reference's src/Frame.cc with the definitions of these two members removed there (or guarded
// by #ifndef ORBX_ADAPTER).  orbxHandle() is exported by adapter/ORBextractor_orbx.cc; orbxVocabulary() returns the device copy
// of the vocabulary, created once from the file System.cc loads.
#include "Frame.h"

#include "orbx_adapter.h"

namespace ORB_SLAM2
{

orbx_extractor* orbxHandle(const ORBextractor* e);
bool orbxVocabularyTransform(const ORBVocabulary* voc, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight);

// replaces Frame.cc:495-669: both pyramids are read where the two extractors left them on the device
void Frame::ComputeStereoMatches()
{
    mvuRight = std::vector<float>(N, -1.0f);
    mvDepth = std::vector<float>(N, -1.0f);
    thread_local orbx_stereo* st = nullptr;
    if (!st && orbxFailed(orbx_stereo_create(&st, 8192, 1, orbxDevice()), "orbx_stereo_create"))
    {
        st = nullptr;
        return;                                            // no depths: the frame is tracked as monocular observations
    }
    int32_t kept = 0;
    // mvKeys / mDescriptors and their right counterparts are exactly what the two extractors returned for this frame
    // (Frame.cc:103-108), so the device reads them where the extractors left them
    if (orbxFailed(orbx_stereo_matches_extractors_host(st, orbxHandle(mpORBextractorLeft), 0, orbxHandle(mpORBextractorRight), 0, N, mbf, mb,
                                                       mvuRight.data(), mvDepth.data(), &kept), "Frame::ComputeStereoMatches"))
    {
        mvuRight.assign(N, -1.0f);
        mvDepth.assign(N, -1.0f);
    }
}

// replaces Frame.cc:286-293: the tree descent of every descriptor runs on the device, the map bookkeeping of
// TemplatedVocabulary::transform (TemplatedVocabulary.h:1160-1200) stays here, in feature order
void Frame::ComputeBoW()
{
    if (!mBowVec.empty())
        return;
    std::vector<int32_t> word(N), node(N);
    std::vector<double> weight(N);
    if (!orbxVocabularyTransform(mpORBvocabulary, mDescriptors.data, N, 4, word.data(), node.data(), weight.data()))
        return;                                            // empty BowVector / FeatureVector, like a frame without features
    for (int i = 0; i < N; i++)
        if (weight[i] > 0)
        {
            mBowVec.addWeight(word[i], weight[i]);
            mFeatVec.addFeature(node[i], i);
        }
    mBowVec.normalize(DBoW2::L1);                    // ORBvoc.txt / ORBvoc.bin: TF_IDF weighting, L1 scoring
}

} // namespace ORB_SLAM2
