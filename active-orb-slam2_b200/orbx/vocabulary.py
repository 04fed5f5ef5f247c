"""Host-side mirror of ORBVocabulary = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB> (reference
Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h) for the one call on the hot path: transform(features, BowVector&,
FeatureVector&, levelsup), used by Frame::ComputeBoW / KeyFrame::ComputeBoW.

The tree descent of every feature runs in liborbx.so (sm_100a CUDA); this module loads a vocabulary file into the arrays
the C ABI takes, and does the map bookkeeping of TemplatedVocabulary.h:1138-1200 on the per-feature results in the
reference's own order (so the BowVector weights are the same doubles: repeated `+=` per hit, then the L1 normalisation)."""
import ctypes as C

import numpy as np

from ._lib import check, lib

TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3          # DBoW2::WeightingType
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = range(6)   # DBoW2::ScoringType


def tree_from_parents(parent, desc, weight, is_leaf, k, L, scoring=L1_NORM, weighting=TF_IDF):
    """node arrays as loadFromBinaryFile / loadFromTextFile build m_nodes: node ids 1..n in file order, children pushed back in
    that order, word ids numbered in the order leaves appear (TemplatedVocabulary.h:1482-1505)."""
    parent = np.asarray(parent, np.int32)
    n = len(parent) + 1
    order = np.argsort(parent, kind="stable")                      # children of a node in ascending id = push_back order
    counts = np.bincount(parent, minlength=n).astype(np.int64)
    child_start = np.zeros(n + 1, np.int32)
    child_start[1:] = np.cumsum(counts)
    children = (order + 1).astype(np.int32)
    node_desc = np.zeros((n, 32), np.uint8)
    node_desc[1:] = desc
    node_weight = np.zeros(n, np.float64)
    node_weight[1:] = np.asarray(weight, np.float32).astype(np.float64)
    word_id = np.full(n, -1, np.int32)
    leaf_nodes = np.nonzero(np.asarray(is_leaf).astype(bool))[0] + 1
    word_id[leaf_nodes] = np.arange(len(leaf_nodes), dtype=np.int32)
    return dict(child_start=child_start, children=children, desc=node_desc, weight=node_weight, word_id=word_id, k=int(k), L=int(L),
                scoring=int(scoring), weighting=int(weighting))


def load_binary(path):
    """TemplatedVocabulary::loadFromBinaryFile (TemplatedVocabulary.h:1467-1508): 24-byte header (nb_nodes, size_node, k, L, scoring,
    weighting), then records of parent (i32), descriptor (32 bytes), weight (f32), is_leaf (u8)."""
    hdr = np.fromfile(path, np.int32, 6)
    nb, size_node, k, L, scoring, weighting = (int(v) for v in hdr)
    if size_node != 41:
        raise ValueError("unexpected node record size %d" % size_node)
    rec = np.dtype([("parent", "<i4"), ("desc", "u1", (32,)), ("weight", "<f4"), ("leaf", "u1")])
    r = np.fromfile(path, rec, offset=24)
    # the reference's read loop runs once more at EOF and re-reads nothing; nb_nodes counts the real records
    r = r[:nb]
    return tree_from_parents(r["parent"], r["desc"], r["weight"], r["leaf"], k, L, scoring, weighting)


class ORBVocabulary:
    def __init__(self, tree, max_features=8192, device=0):
        self._L = lib()
        self._h = C.c_void_p()
        self.tree = tree
        self._keep = [np.ascontiguousarray(tree[k], t) for k, t in (("child_start", np.int32), ("children", np.int32), ("desc", np.uint8),
                                                                    ("weight", np.float64), ("word_id", np.int32))]
        cs, ch, nd, w, wid = self._keep
        check(self._L.orbx_vocabulary_create(C.byref(self._h), len(cs) - 1, cs.ctypes.data, ch.ctypes.data, nd.ctypes.data, w.ctypes.data,
                                             wid.ctypes.data, tree["L"], max_features, device))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.orbx_vocabulary_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def transform_features(self, features, levelsup=4):
        """-> (word[n], node[n], weight[n]) of every feature (the device part)"""
        d = np.ascontiguousarray(features, np.uint8).reshape(-1, 32)
        n = len(d)
        word, node, wt = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.float64)
        check(self._L.orbx_vocabulary_transform_host(self._h, d.ctypes.data, n, levelsup, word.ctypes.data, node.ctypes.data, wt.ctypes.data))
        return word[:n], node[:n], wt[:n]

    def transform(self, features, levelsup=4):
        """transform(features, BowVector &v, FeatureVector &fv, levelsup) -> (v, fv) as dicts ordered like the std::maps"""
        word, node, wt = self.transform_features(features, levelsup)
        return bow_maps(word, node, wt, self.tree["weighting"], self.tree["scoring"])

    def transform_device(self, levelsup, d_desc, d_counts, count_step, pitch, max_count, batch, d_word, d_node, d_weight, stream=0):
        check(self._L.orbx_vocabulary_transform_device(self._h, levelsup, d_desc, d_counts, count_step, pitch, max_count, batch, d_word, d_node,
                                                       d_weight, stream))

    def last_launches(self):
        return self._L.orbx_vocabulary_last_launches(self._h)


def bow_maps(word, node, weight, weighting=TF_IDF, scoring=L1_NORM):
    """the caller-side loop of TemplatedVocabulary::transform (TemplatedVocabulary.h:1138-1200) on per-feature results"""
    v, fv = {}, {}
    must_l1 = scoring in (L1_NORM, CHI_SQUARE, KL, BHATTACHARYYA)        # mustNormalize(): these scorings ask for the L1 norm
    must_l2 = scoring == L2_NORM
    for i, (wid, nid, w) in enumerate(zip(word.tolist(), node.tolist(), weight.tolist())):
        if w > 0:                                                          # not stopped
            if weighting in (TF, TF_IDF):
                v[wid] = v[wid] + w if wid in v else w                     # BowVector::addWeight
            elif wid not in v:
                v[wid] = w                                                 # addIfNotExist
            fv.setdefault(nid, []).append(i)                               # FeatureVector::addFeature
    v = dict(sorted(v.items()))
    fv = dict(sorted(fv.items()))
    if weighting in (TF, TF_IDF) and v and not (must_l1 or must_l2):
        nd = float(len(v))
        v = {k: x / nd for k, x in v.items()}
    if must_l1 or must_l2:                                                 # BowVector::normalize, BowVector.cpp:59-86
        norm = 0.0
        for x in v.values():
            norm += abs(x) if must_l1 else x * x
        if must_l2:
            norm = float(np.sqrt(norm))
        if norm > 0.0:
            v = {k: x / norm for k, x in v.items()}
    return v, fv


def feature_vector_csr(fv):
    """FeatureVector dict -> (node ids ascending, CSR starts, feature indices): the form orbx_bow_set takes"""
    ids = np.array(list(fv.keys()), np.uint32)
    start = np.zeros(len(ids) + 1, np.int32)
    feat = []
    for j, k in enumerate(fv):
        feat += fv[k]
        start[j + 1] = len(feat)
    return ids, start, np.array(feat, np.int32)
