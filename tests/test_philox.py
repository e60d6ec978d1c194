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
