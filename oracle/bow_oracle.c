/* orbx CPU oracle, bag-of-words transform — TEST INFRASTRUCTURE ONLY (see orbx_oracle.h).
 *
 * Restates the per-feature tree descent of DBoW2 (reference Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1216-1262,
 * TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)) with FORB::distance (FORB.cpp:98-117, the
 * 256-bit Hamming distance): from the root, move to the child with the smallest distance (the FIRST one among equals:
 * `d < best_d`), remember the node passed at level L - levelsup, stop at a leaf; return its word id and weight.
 * The caller-side loop (TemplatedVocabulary.h:1138-1200: addWeight / addFeature per feature, L1 normalisation) is map
 * bookkeeping on these per-feature results and is done by the host adapter in the reference's own order.
 * PARITY PINNING: the reference has no tests.  PINNED against the reference's own ORBVocabulary (DBoW2 TemplatedVocabulary<FORB>,
 * its loadFromTextFile and transform, compiled into oracle/_ref/liborbmatcher_ref.so): same word per feature, bit-identical
 * BowVector, same FeatureVector for every weighting / scoring pair tried (tests/test_oracle_ref_matcher.py).  One defined choice:
 * for a branch that ends above level L - levelsup the reference never writes `nid` (an uninitialised NodeId in the caller,
 * TemplatedVocabulary.h:1160-1170); oracle and kernel return node 0.  tests/test_bow_oracle.py also checks this file against an
 * independent numpy statement on synthetic trees and on the reference's own Vocabulary/ORBvoc.bin.
 */
#include "orbx_oracle.h"
#include <stddef.h>

void orbo_bow_transform(const orbo_vocabulary *V, const uint8_t *desc, int n, int levelsup, int32_t *word, int32_t *node,
                        double *weight) {
    const int nid_level = V->L - levelsup;
    for (int i = 0; i < n; i++) {
        const uint8_t *f = desc + (size_t)32 * i;
        int32_t nid = 0, final_id = 0;                        /* `if(nid_level <= 0) *nid = 0` */
        int current_level = 0;
        if (V->child_start[1] == V->child_start[0]) {         /* a vocabulary without nodes below the root: empty() */
            word[i] = -1; node[i] = 0; weight[i] = 0;
            continue;
        }
        do {
            ++current_level;
            const int c0 = V->child_start[final_id], c1 = V->child_start[final_id + 1];
            int best = V->children[c0];
            int best_d = orbo_hamming256(f, V->desc + (size_t)32 * best);
            for (int c = c0 + 1; c < c1; c++) {
                const int id = V->children[c];
                const int d = orbo_hamming256(f, V->desc + (size_t)32 * id);
                if (d < best_d) { best_d = d; best = id; }
            }
            final_id = best;
            if (current_level == nid_level) nid = final_id;
        } while (V->child_start[final_id + 1] > V->child_start[final_id]);   /* !isLeaf() */
        word[i] = V->word_id[final_id];
        weight[i] = V->weight[final_id];
        node[i] = nid;
    }
}
