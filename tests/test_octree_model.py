"""The rank/round formulation of DistributeOctTree (tests/octree_model.py, mirrored by csrc/octree.cu) must give
the oracle's (== reference ORBextractor.cc:539-763) keypoints in the same order.  CPU only."""
import numpy as np
import pytest

from oracle import oracle_py as O
import octree_model as M


def random_candidates(rng, width, height, n, clustered):
    pts = set()
    while len(pts) < n:
        if clustered and rng.random() < 0.7:
            cx, cy = rng.integers(3, width - 3), rng.integers(3, height - 3)
            x = int(np.clip(cx + rng.integers(-6, 7), 3, width - 4))
            y = int(np.clip(cy + rng.integers(-6, 7), 3, height - 4))
        else:
            x, y = int(rng.integers(3, width - 3)), int(rng.integers(3, height - 3))
        pts.add((x, y))
    pts = list(pts)
    rng.shuffle(pts)
    return pts


@pytest.mark.parametrize("geom", [(608, 448, 217), (147, 102, 60), (1209, 344, 434), (314, 73, 122), (608, 448, 0),
                                   (608, 448, 5), (150, 200, 30), (300, 100, 1000)])
@pytest.mark.parametrize("n", [0, 1, 2, 7, 60, 400, 3000])
def test_model_matches_oracle(geom, n):
    width, height, N = geom
    rng = np.random.default_rng(n * 7919 + width)
    for trial in range(3):
        n_eff = min(n, (width - 6) * (height - 6) // 3)
        pts = random_candidates(rng, width, height, n_eff, clustered=trial == 1)
        sc = rng.integers(7, 60 if trial == 2 else 255, len(pts))
        K = np.zeros(len(pts), O.KP_DTYPE)
        K["x"] = [p[0] for p in pts]
        K["y"] = [p[1] for p in pts]
        K["response"] = sc
        ref = O.distribute(K, 16, 16 + width, 16, 16 + height, N)
        got = M.distribute([p[0] for p in pts], [p[1] for p in pts], sc.tolist(), list(range(len(pts))), width, height, N)
        assert len(ref) == len(got)
        assert np.array_equal(ref["x"], K["x"][got]) and np.array_equal(ref["y"], K["y"][got])
        assert np.array_equal(ref["response"], K["response"][got])
