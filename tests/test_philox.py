"""Philox4x32-10: the Python restatement used by the tests is pinned to the Random123 known-answer
vectors on CPU; the device implementation is pinned to the same vectors on the GPU."""
import pytest

from tests.helpers import PHILOX_KAT, philox4x32_10


def test_philox_kat_cpu():
    for ctr, key, want in PHILOX_KAT:
        assert tuple(philox4x32_10(ctr, key)) == want


@pytest.mark.gpu
def test_philox_kat_gpu():
    from ndsimulator_b200.engine import philox_kat
    for ctr, key, want in PHILOX_KAT:
        assert tuple(philox_kat(ctr, key)) == want


def test_reference_normal_stream_is_standard_normal():
    """The six-normals-per-call layout (21-bit Box-Muller fields) restated in tests/helpers.py: N(0,1)
    marginals, uncorrelated lanes, tails cut at sqrt(2 ln 2^21) = 5.40 sigma."""
    import numpy as np
    from scipy import stats
    from tests.helpers import philox_normals_f64
    z = np.array([philox_normals_f64(2024, w, 60) for w in range(400)])     # 400 walkers x 10 calls x 6 lanes
    assert stats.kstest(z.reshape(-1), "norm").pvalue > 1e-3
    lanes = z.reshape(400, 10, 6)
    for a in range(6):
        assert abs(lanes[:, :, a].mean()) < 5 / np.sqrt(4000)
        assert abs(lanes[:, :, a].var() - 1) < 0.12
        for b in range(a + 1, 6):
            assert abs(np.corrcoef(lanes[:, :, a].ravel(), lanes[:, :, b].ravel())[0, 1]) < 0.08
    assert np.abs(z).max() <= np.sqrt(2 * np.log(2.0 ** 21)) + 1e-12
