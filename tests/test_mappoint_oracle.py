"""Oracle of MapPoint::ComputeDistinctiveDescriptors (oracle/mappoint_oracle.c, reference src/MapPoint.cc:275-340) against numpy."""
import numpy as np

from oracle import oracle_py as O

POP = np.array([bin(i).count("1") for i in range(256)], np.int32)


def observation_sets(seed, n_points=200, max_obs=40):
    rng = np.random.default_rng(seed)
    sets = []
    for p in range(n_points):
        n = int(rng.integers(0, max_obs + 1)) if p % 7 else int(rng.integers(0, 3))
        base = rng.integers(0, 256, 32).astype(np.uint8)
        d = np.tile(base, (n, 1))
        for i in range(n):                                   # observations of one point: the same patch seen again, some bits off
            for b in rng.choice(256, int(rng.integers(0, 60)), replace=False):
                d[i, b >> 3] ^= np.uint8(1 << (b & 7))
        if n > 3 and p % 5 == 0:
            d[2] = d[0]                                      # duplicates: equal medians, the first row must win
        sets.append(d)
    return sets


def numpy_best(d):
    n = len(d)
    if n == 0:
        return -1, -1
    M = POP[np.bitwise_xor(d[:, None, :], d[None, :, :])].sum(2)
    med = np.sort(M, 1)[:, int(0.5 * (n - 1))]
    i = int(np.argmin(med))                                  # first minimum
    return i, int(med[i])


def to_csr(sets):
    start = np.zeros(len(sets) + 1, np.int32)
    start[1:] = np.cumsum([len(s) for s in sets])
    desc = np.concatenate(sets) if start[-1] else np.zeros((0, 32), np.uint8)
    return start, desc


def test_oracle_equals_numpy():
    for seed in range(3):
        sets = observation_sets(seed)
        bi, bm = O.distinctive_descriptors(*to_csr(sets))
        ref = [numpy_best(s) for s in sets]
        assert bi.tolist() == [r[0] for r in ref] and bm.tolist() == [r[1] for r in ref]
        assert (bi == -1).sum() > 0 and (bi > 0).sum() > 50
