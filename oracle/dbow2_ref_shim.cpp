// TEST INFRASTRUCTURE ONLY: a C entry point over the REFERENCE's own DBoW2::BowVector / DBoW2::FeatureVector.
// oracle/Makefile (target `ref`) compiles this file together with /root/reference/Thirdparty/DBoW2/DBoW2/BowVector.cpp and
// FeatureVector.cpp -- from where they lie, nothing is copied -- into oracle/_ref/libdbow2_ref.so.  These two classes are the
// only part of the reference's hot path that builds in this image (pure STL; everything else needs OpenCV, Eigen or g2o).
// The function replays the caller-side loop of TemplatedVocabulary::transform (TemplatedVocabulary.h:1138-1200) on per-feature
// (word id, node id, weight) triples with the reference's addWeight / addIfNotExist / addFeature / normalize, so that the
// product's host bookkeeping (orbx/vocabulary.py: bow_maps) is pinned against reference code that actually ran.
#include "BowVector.h"
#include "FeatureVector.h"

#include <cstdint>

extern "C" int dbow2_ref_maps(int n, const uint32_t *word, const uint32_t *node, const double *weight, int tf_like /* TF or TF_IDF */,
                              int must_normalize, int l2, uint32_t *v_ids, double *v_vals, int *n_v, uint32_t *fv_ids,
                              int32_t *fv_start, uint32_t *fv_feat, int *n_fv) {
    DBoW2::BowVector v;
    DBoW2::FeatureVector fv;
    for (int i = 0; i < n; i++) {
        if (weight[i] > 0) {                                  // not stopped
            if (tf_like) v.addWeight(word[i], weight[i]);
            else v.addIfNotExist(word[i], weight[i]);
            fv.addFeature(node[i], (unsigned int)i);
        }
    }
    if (tf_like && !v.empty() && !must_normalize) {
        const double nd = v.size();
        for (DBoW2::BowVector::iterator vit = v.begin(); vit != v.end(); vit++) vit->second /= nd;
    }
    if (must_normalize) v.normalize(l2 ? DBoW2::L2 : DBoW2::L1);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = v.begin(); it != v.end(); ++it, ++k) { v_ids[k] = it->first; v_vals[k] = it->second; }
    *n_v = k;
    k = 0;
    int f = 0;
    fv_start[0] = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++k) {
        fv_ids[k] = it->first;
        for (size_t j = 0; j < it->second.size(); j++) fv_feat[f++] = it->second[j];
        fv_start[k + 1] = f;
    }
    *n_fv = k;
    return 0;
}
