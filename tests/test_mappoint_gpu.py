"""GPU parity of MapPoint::ComputeDistinctiveDescriptors (orbx_mappoints_* through the C ABI) against the CPU oracle: identical
index (and median) for every map point."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx.mappoint import MapPointOps
from test_mappoint_oracle import observation_sets, to_csr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n_points,max_obs", [(0, 200, 40), (1, 1500, 25), (2, 30, 150), (3, 1, 1)])
def test_matches_oracle(seed, n_points, max_obs):
    sets = observation_sets(seed, n_points, max_obs)
    ops = MapPointOps(max_points=2048, max_descriptors=65536)
    best, med = ops.ComputeDistinctiveDescriptors(sets)
    rb, rm = O.distinctive_descriptors(*to_csr(sets))
    assert np.array_equal(best, rb) and np.array_equal(med, rm)
    assert ops.last_launches() == 1
    ops.close()


def test_empty_and_capacity():
    ops = MapPointOps(max_points=4, max_descriptors=10)
    best, med = ops.ComputeDistinctiveDescriptors([])
    assert len(best) == 0
    best, med = ops.ComputeDistinctiveDescriptors([np.zeros((0, 32), np.uint8), np.zeros((2, 32), np.uint8)])
    assert best.tolist() == [-1, 0] and med.tolist() == [-1, 0]
    from orbx._lib import OrbxError
    with pytest.raises(OrbxError):
        ops.ComputeDistinctiveDescriptors([np.zeros((11, 32), np.uint8)])
    with pytest.raises(OrbxError):
        ops.ComputeDistinctiveDescriptors([np.zeros((1, 32), np.uint8)] * 5)
    ops.close()
