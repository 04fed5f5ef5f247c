"""Oracle of the DBoW2 tree descent (oracle/bow_oracle.c, reference TemplatedVocabulary.h:1216-1262) against an independent
numpy statement, on synthetic trees and -- where the reference tree is mounted -- on its own Vocabulary/ORBvoc.bin; plus
the host-side map bookkeeping (orbx.vocabulary.bow_maps, TemplatedVocabulary.h:1138-1200).  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth
from orbx.vocabulary import bow_maps, feature_vector_csr, load_binary, tree_from_parents

ORBVOC = "/root/reference/Vocabulary/ORBvoc.bin"
POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def descend_numpy(tree, f, levelsup):
    cs, ch = tree["child_start"], tree["children"]
    nid_level, cur, nid, level = tree["L"] - levelsup, 0, 0, 0
    if cs[1] == cs[0]:
        return -1, 0, 0.0
    while True:
        level += 1
        kids = ch[cs[cur]:cs[cur + 1]]
        d = POP[np.bitwise_xor(tree["desc"][kids], f)].sum(1)
        cur = int(kids[int(np.argmin(d))])                 # argmin returns the first minimum
        if level == nid_level:
            nid = cur
        if cs[cur + 1] == cs[cur]:
            return int(tree["word_id"][cur]), nid, float(tree["weight"][cur])


@pytest.mark.parametrize("seed,k,L,levelsup", [(0, 10, 4, 2), (1, 10, 3, 4), (2, 4, 6, 4), (3, 10, 5, 1), (4, 2, 7, 3), (5, 10, 4, 0)])
def test_oracle_equals_numpy_on_synthetic_trees(seed, k, L, levelsup):
    tree = tree_from_parents(*synth.random_vocabulary(seed, k=k, L=L))
    rng = np.random.default_rng(seed)
    feats = np.concatenate([synth.descriptors_near_words(rng, tree, 300), rng.integers(0, 256, (100, 32)).astype(np.uint8)])
    word, node, wt = O.bow_transform(tree, feats, levelsup)
    ref = [descend_numpy(tree, f, levelsup) for f in feats]
    assert word.tolist() == [r[0] for r in ref]
    assert node.tolist() == [r[1] for r in ref]
    assert wt.tolist() == [r[2] for r in ref]
    assert (word >= 0).all() and len(set(word.tolist())) > 20


def test_ties_keep_the_first_child():
    # two children with the same descriptor: the first one (lower position in the child list) must win
    parent = np.array([0, 0, 0], np.int32)
    desc = np.zeros((3, 32), np.uint8); desc[2] = 255
    tree = tree_from_parents(parent, desc, [1.0, 2.0, 3.0], [1, 1, 1], 3, 1)
    word, node, wt = O.bow_transform(tree, np.zeros((1, 32), np.uint8), 0)
    assert word[0] == 0 and wt[0] == 1.0 and node[0] == 1        # L - levelsup = 1: the node at level 1 is the leaf itself


def test_bookkeeping_matches_the_reference_loop():
    tree = tree_from_parents(*synth.random_vocabulary(7, k=6, L=3, stop=0.2))
    rng = np.random.default_rng(7)
    feats = synth.descriptors_near_words(rng, tree, 500, flip=5)
    word, node, wt = O.bow_transform(tree, feats, 1)
    v, fv = bow_maps(word, node, wt)
    assert list(v) == sorted(v) and list(fv) == sorted(fv)
    assert abs(sum(v.values()) - 1.0) < 1e-12                      # L1-normalised
    live = wt > 0
    assert sorted(sum(fv.values(), [])) == np.nonzero(live)[0].tolist()     # stopped words contribute no feature
    for nid, idx in fv.items():
        assert idx == sorted(idx) and (node[idx] == nid).all()
    # a word hit c times carries c additions of its weight before normalisation
    wid = max(v, key=lambda k_: (word == k_).sum())
    c, w = int((word == wid).sum()), float(wt[word == wid][0])
    acc = 0.0
    for _ in range(c):
        acc += w
    total = 0.0
    for k_ in sorted(set(word[live].tolist())):
        s = 0.0
        for _ in range(int((word == k_).sum())):
            s += float(wt[word == k_][0])
        total += abs(s)
    assert v[wid] == acc / total
    ids, start, feat = feature_vector_csr(fv)
    assert ids.tolist() == list(fv) and feat.tolist() == sum(fv.values(), []) and start[-1] == len(feat)


def test_empty_inputs():
    tree = tree_from_parents(*synth.random_vocabulary(1, k=3, L=2))
    word, node, wt = O.bow_transform(tree, np.zeros((0, 32), np.uint8), 4)
    assert len(word) == 0
    assert bow_maps(word, node, wt) == ({}, {})


@pytest.mark.skipif(not os.path.exists(ORBVOC), reason="the reference's vocabulary file is only mounted in the build container")
def test_oracle_on_the_reference_vocabulary():
    tree = load_binary(ORBVOC)
    assert (tree["k"], tree["L"]) == (10, 6) and len(tree["child_start"]) - 1 == 1082074 and (tree["word_id"] >= 0).sum() > 900000
    assert tree["weighting"] == 0 and tree["scoring"] == 0           # TF_IDF, L1_NORM
    rng = np.random.default_rng(0)
    feats = np.concatenate([synth.descriptors_near_words(rng, tree, 150, flip=30), rng.integers(0, 256, (50, 32)).astype(np.uint8)])
    word, node, wt = O.bow_transform(tree, feats, 4)
    ref = [descend_numpy(tree, f, 4) for f in feats]
    assert word.tolist() == [r[0] for r in ref] and node.tolist() == [r[1] for r in ref] and wt.tolist() == [r[2] for r in ref]
    # level-2 nodes (L - levelsup = 2): at most k + k^2 distinct ids, all children of level-1 nodes
    lvl1 = set(tree["children"][tree["child_start"][0]:tree["child_start"][1]].tolist())
    par = np.zeros(len(tree["child_start"]) - 1, np.int64)
    for p in lvl1:
        par[tree["children"][tree["child_start"][p]:tree["child_start"][p + 1]]] = p
    assert all(int(par[n]) in lvl1 for n in node.tolist())


REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdbow2_ref.so")


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libdbow2_ref.so is built from the reference tree (make -C oracle ref)")
@pytest.mark.parametrize("weighting,scoring", [(0, 0), (1, 0), (0, 1), (0, 5), (2, 0), (3, 5)])
def test_bookkeeping_against_the_reference_classes(weighting, scoring):
    """bow_maps (the product's host bookkeeping) against the REFERENCE's compiled DBoW2::BowVector / FeatureVector: identical keys,
    identical doubles bit for bit, identical feature lists -- for every weighting / normalisation branch of transform()"""
    import ctypes as C
    L = C.CDLL(REF_SO)
    rng = np.random.default_rng(weighting * 7 + scoring)
    n = 3000
    word = rng.integers(0, 400, n).astype(np.uint32)
    wts = rng.uniform(0.01, 9.0, 400)
    wts[rng.random(400) < 0.1] = 0.0                                   # stopped words
    weight = wts[word].astype(np.float64)
    node = (word // 7).astype(np.uint32)
    v, fv = bow_maps(word.astype(np.int64), node.astype(np.int64), weight, weighting, scoring)
    v_ids, v_vals = np.zeros(n, np.uint32), np.zeros(n, np.float64)
    fv_ids, fv_start, fv_feat = np.zeros(n, np.uint32), np.zeros(n + 1, np.int32), np.zeros(n, np.uint32)
    n_v, n_fv = C.c_int(), C.c_int()
    must = scoring in (0, 1, 2, 3, 4)
    L.dbow2_ref_maps(n, word.ctypes.data_as(C.c_void_p), node.ctypes.data_as(C.c_void_p), weight.ctypes.data_as(C.c_void_p),
                     int(weighting in (0, 1)), int(must), int(scoring == 1), v_ids.ctypes.data_as(C.c_void_p), v_vals.ctypes.data_as(C.c_void_p),
                     C.byref(n_v), fv_ids.ctypes.data_as(C.c_void_p), fv_start.ctypes.data_as(C.c_void_p), fv_feat.ctypes.data_as(C.c_void_p),
                     C.byref(n_fv))
    assert list(v) == v_ids[:n_v.value].tolist()
    assert np.array(list(v.values()), np.float64).tobytes() == v_vals[:n_v.value].tobytes()
    assert list(fv) == fv_ids[:n_fv.value].tolist()
    for k, nid in enumerate(fv):
        assert fv[nid] == fv_feat[fv_start[k]:fv_start[k + 1]].tolist()
    assert n_v.value > 100 and n_fv.value > 20
