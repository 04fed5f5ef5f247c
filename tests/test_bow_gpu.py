"""GPU parity of the DBoW2 transform (liborbx.so, orbx_vocabulary_* through the C ABI) against the CPU oracle.
Bar: identical word ids, node ids and weights per feature, hence identical BowVector / FeatureVector maps."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth
from orbx.vocabulary import ORBVocabulary, bow_maps, tree_from_parents

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,k,L,levelsup,n", [(0, 10, 4, 2, 1000), (1, 10, 3, 4, 7), (2, 4, 6, 4, 333), (3, 25, 3, 1, 500), (4, 2, 7, 3, 64),
                                                 (5, 10, 5, 0, 2000), (6, 10, 6, 4, 1500)])
def test_transform_matches_oracle(seed, k, L, levelsup, n):
    tree = tree_from_parents(*synth.random_vocabulary(seed, k=k, L=L, prune=0.25 if L > 5 else 0.1))
    rng = np.random.default_rng(seed)
    feats = np.concatenate([synth.descriptors_near_words(rng, tree, n - n // 4), rng.integers(0, 256, (n // 4, 32)).astype(np.uint8)])
    voc = ORBVocabulary(tree, max_features=4096)
    word, node, wt = voc.transform_features(feats, levelsup)
    rw, rn, rwt = O.bow_transform(tree, feats, levelsup)
    assert np.array_equal(word, rw) and np.array_equal(node, rn) and np.array_equal(wt, rwt)
    assert voc.last_launches() == 1
    v, fv = voc.transform(feats, levelsup)
    rv, rfv = bow_maps(rw, rn, rwt)
    assert v == rv and fv == rfv and list(v) == list(rv)
    voc.close()


def test_ties_empty_and_capacity():
    parent = np.array([0, 0, 0], np.int32)
    desc = np.zeros((3, 32), np.uint8); desc[2] = 255
    tree = tree_from_parents(parent, desc, [1.0, 2.0, 3.0], [1, 1, 1], 3, 1)
    voc = ORBVocabulary(tree, max_features=16)
    word, node, wt = voc.transform_features(np.zeros((5, 32), np.uint8), 0)
    assert (word == 0).all() and (wt == 1.0).all() and (node == 1).all()
    word, node, wt = voc.transform_features(np.zeros((0, 32), np.uint8), 0)
    assert len(word) == 0
    from orbx._lib import OrbxError
    with pytest.raises(OrbxError):
        voc.transform_features(np.zeros((17, 32), np.uint8), 0)
    voc.close()
    with pytest.raises(OrbxError):                                     # child id outside the tree
        bad = dict(tree); bad["children"] = np.array([1, 2, 9], np.int32)
        ORBVocabulary(bad)


def test_batched_device_frames():
    """descriptors where an extractor would leave them: [frame][pitch][32] + counts on the device"""
    import torch
    tree = tree_from_parents(*synth.random_vocabulary(11, k=10, L=5))
    voc = ORBVocabulary(tree)
    rng = np.random.default_rng(11)
    B, pitch = 6, 700
    counts = np.array([700, 0, 1, 333, 699, 32], np.int32)
    desc = np.zeros((B, pitch, 32), np.uint8)
    for f in range(B):
        desc[f, :counts[f]] = synth.descriptors_near_words(rng, tree, counts[f]) if counts[f] else 0
    d_desc, d_cnt = torch.from_numpy(desc).cuda(), torch.from_numpy(counts).cuda()
    d_word = torch.full((B, pitch), -7, dtype=torch.int32, device="cuda")
    d_node = torch.full((B, pitch), -7, dtype=torch.int32, device="cuda")
    d_wt = torch.zeros((B, pitch), dtype=torch.float64, device="cuda")
    voc.transform_device(4, d_desc.data_ptr(), d_cnt.data_ptr(), 1, pitch, pitch, B, d_word.data_ptr(), d_node.data_ptr(), d_wt.data_ptr(),
                         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    word, node, wt = d_word.cpu().numpy(), d_node.cpu().numpy(), d_wt.cpu().numpy()
    for f in range(B):
        n = counts[f]
        rw, rn, rwt = O.bow_transform(tree, desc[f, :n], 4)
        assert np.array_equal(word[f, :n], rw) and np.array_equal(node[f, :n], rn) and np.array_equal(wt[f, :n], rwt), f
        assert (word[f, n:] == -7).all()                               # nothing written past the frame's count
    voc.close()
