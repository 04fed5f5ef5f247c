// Device copy of an ORBVocabulary (DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) for orbx_vocabulary_transform_*:
// m_nodes flattened once into the arrays orbx_vocabulary_create takes (children CSR in m_nodes[id].children order, 32-byte
// descriptors, weights, word ids).  The vocabulary class itself is untouched; its protected members are read through a
// derived accessor.  One device copy per vocabulary object, made on first use (System.cc loads the vocabulary once).
#include "ORBVocabulary.h"

#include "orbx_adapter.h"

#include <map>
#include <mutex>
#include <vector>

namespace ORB_SLAM2
{

namespace
{
struct VocabularyAccess : public ORBVocabulary
{
    // Node is a protected nested type, so the flattening is a member
    orbx_vocabulary* toDevice() const
    {
        const int n = (int)m_nodes.size();
        std::vector<int32_t> childStart(n + 1, 0), children, wordId(n, -1);
        std::vector<uint8_t> desc((size_t)32 * n, 0);
        std::vector<double> weight(n, 0.0);
        for (int i = 0; i < n; i++)
        {
            const Node& nd = m_nodes[i];
            childStart[i] = (int32_t)children.size();
            for (size_t c = 0; c < nd.children.size(); c++)
                children.push_back((int32_t)nd.children[c]);
            if (i > 0 && nd.descriptor.data)
                for (int k = 0; k < 32; k++) desc[(size_t)32 * i + k] = nd.descriptor.data[k];
            weight[i] = nd.weight;
            if (nd.isLeaf() && i > 0)
                wordId[i] = (int32_t)nd.word_id;
        }
        childStart[n] = (int32_t)children.size();
        orbx_vocabulary* h = nullptr;
        if (orbxFailed(orbx_vocabulary_create(&h, n, childStart.data(), children.data(), desc.data(), weight.data(), wordId.data(), m_L, 8192,
                                              orbxDevice()), "orbx_vocabulary_create"))
            return nullptr;
        return h;
    }
};

std::mutex gMutex;
std::map<const ORBVocabulary*, orbx_vocabulary*> gDevice;
} // namespace

orbx_vocabulary* orbxVocabulary(const ORBVocabulary* voc)
{
    std::lock_guard<std::mutex> lock(gMutex);
    std::map<const ORBVocabulary*, orbx_vocabulary*>::iterator it = gDevice.find(voc);
    if (it != gDevice.end())
        return it->second;
    orbx_vocabulary* h = static_cast<const VocabularyAccess*>(voc)->toDevice();
    gDevice[voc] = h;
    return h;
}

// orbx_vocabulary_transform_host stages through the handle's own buffers and stream (include/orbx.h: handles are not re-entrant), and
// one device copy serves every caller of a vocabulary object: Frame::ComputeBoW on the Tracking thread, KeyFrame::ComputeBoW on the
// LocalMapping thread once it is moved over.  The calls are serialised here; the tree arrays are shared and read-only.
bool orbxVocabularyTransform(const ORBVocabulary* voc, const uint8_t* desc, int n, int levelsup, int32_t* word, int32_t* node, double* weight)
{
    orbx_vocabulary* h = orbxVocabulary(voc);
    if (!h)
        return false;
    static std::mutex gTransformMutex;
    std::lock_guard<std::mutex> lock(gTransformMutex);
    return !orbxFailed(orbx_vocabulary_transform_host(h, desc, n, levelsup, word, node, weight), "orbx_vocabulary_transform_host");
}

} // namespace ORB_SLAM2
