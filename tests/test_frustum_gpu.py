"""GPU parity of Frame::isInFrustum (orbx_frustum_* through the C ABI) against the CPU oracle: every field of every record
bit-equal; the level of the records the device flags as undecided is settled on the host with libm's logf."""
import numpy as np
import pytest

from oracle import oracle_py as O
from orbx import synth
from orbx.frustum import isInFrustum
from test_frustum_oracle import pack

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n", [(0, 4000), (1, 1), (2, 20000), (3, 777)])
def test_matches_oracle(seed, n):
    fr, p = pack(*synth.frustum_scene(seed, n=n))
    ref = O.is_in_frustum(fr, p)
    raw, n_amb = isInFrustum(fr, p, resolve=False)
    clear = raw["pad"][:, 0] == 0
    assert raw[clear].tobytes() == ref[clear].tobytes()                       # decided on the device: identical records
    assert n_amb == int((~clear).sum()) and n_amb <= max(2, n // 500)         # the undecided ones are rare
    got, _ = isInFrustum(fr, p)
    assert got.tobytes() == ref.tobytes()


def test_level_boundaries_are_flagged_not_guessed():
    """points whose distance ratio is a power of the scale factor: logf decides the level by its last bit"""
    fr, p = pack(*synth.frustum_scene(5, n=64))
    lsf, sf = fr["log_scale_factor"], np.float32(1.2)
    p["skip"] = 0
    # put every point straight ahead of the camera at distance d, normal facing the camera, max_distance = d * sf^k
    R = fr["Rcw"].reshape(3, 3).astype(np.float64)
    for i in range(len(p)):
        d = 2.0 + 0.05 * i
        Pw = R.T @ (np.array([0, 0, d]) - fr["tcw"].astype(np.float64))
        p["x"][i], p["y"][i], p["z"][i] = Pw.astype(np.float32)
        n = (Pw - fr["Ow"].astype(np.float64)); n /= np.linalg.norm(n)
        p["nx"][i], p["ny"][i], p["nz"][i] = n.astype(np.float32)
        o = np.array([p["x"][i], p["y"][i], p["z"][i]], np.float32) - fr["Ow"]
        dist = np.float32(np.sqrt(np.sum(o.astype(np.float64) ** 2)))
        p["max_distance"][i] = dist * sf ** np.float32(i % 7)
        p["min_distance"][i] = p["max_distance"][i] / np.float32(4.0)
    ref = O.is_in_frustum(fr, p)
    raw, n_amb = isInFrustum(fr, p, resolve=False)
    assert ref["in_view"].sum() > 40 and n_amb > 20                           # most of them sit on a boundary
    got, _ = isInFrustum(fr, p)
    assert got.tobytes() == ref.tobytes()


def test_empty():
    fr, p = pack(*synth.frustum_scene(0, n=8))
    got, n_amb = isInFrustum(fr, p[:0])
    assert len(got) == 0 and n_amb == 0
