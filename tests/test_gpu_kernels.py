"""Kernel-level checks that do not need the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sqrt_matches_ieee(pkg):
    """The solver's branch-free sqrt (dymu_sqrt_normal) must equal IEEE sqrt bit for bit."""
    dev = pkg.cuda_api.DeviceLayer(64, 64)
    for seed in (1, 2, 3):
        bad, first = dev.selftest_sqrt(200_000_000, seed)
        assert bad == 0, "first mismatch at x=%r" % first


def test_uniform_cost_known_answer(pkg):
    """Uniform cost c: along the goal's row and column T = c * distance exactly."""
    n, c = 97, 1.5
    dev = pkg.cuda_api.DeviceLayer(n, n)
    dev.set_cost_map(np.full((n, n), c))
    dev.solve_total_cost([(48, 48)])
    T = dev.download_total_cost()
    assert np.array_equal(T[48, 48:], c * np.arange(0, n - 48))
    assert np.array_equal(T[:49, 48][::-1], c * np.arange(0, 49))
    assert np.all(np.isfinite(T))


def test_wall_blocks_propagation(pkg):
    n = 80
    cost = np.ones((n, n))
    cost[:, 40] = 0.0          # full-height obstacle wall
    dev = pkg.cuda_api.DeviceLayer(n, n)
    dev.set_cost_map(cost)
    dev.solve_total_cost([(10, 10)])
    T = dev.download_total_cost(xform=pkg.cuda_api.XFORM_INF_TO_MINUS1)
    assert np.all(T[:, 40:] == -1.0) and np.all(T[:, :40] >= 0.0)
