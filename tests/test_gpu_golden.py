"""B200 library against the committed golden vectors (outputs of the unmodified reference)."""
import numpy as np
import pytest

import golden_cases as gc
from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_config1_golden(pkg):
    g, o = gc.replay_config1(pkg.DyMuPathPlanner)
    assert np.array_equal(o["obstacle"], g["obstacle"])
    assert rel_err(o["cost"], g["cost"]) <= 1e-14
    closed = g["state_early"] > 0
    assert np.array_equal(o["state_early"] > 0, closed)
    assert rel_err(np.where(closed, o["total_cost_early"], 0), np.where(closed, g["total_cost_early"], 0)) <= 1e-9
    assert np.array_equal(o["total_cost_full"] < 0, g["total_cost_full"] < 0)
    assert rel_err(o["total_cost_full"], g["total_cost_full"]) <= 1e-9
    assert o["path_full"].shape == g["path_full"].shape
    assert np.max(np.abs(o["path_full"][:, :2] - g["path_full"][:, :2])) <= 1e-3
    assert o["path_early"].shape == g["path_early"].shape
    assert np.max(np.abs(o["path_early"][:, :2] - g["path_early"][:, :2])) <= 1e-3


@pytest.mark.parametrize("approach,name", [(1, "repair_120_sweeping.npz"),
                                           (0, "repair_120_conservative.npz")])
def test_repair_golden(pkg, approach, name):
    g, o = gc.replay_repair(pkg.DyMuPathPlanner, approach, name)
    assert o["repaired"]
    assert np.array_equal(o["risk"] > 0, g["risk"] > 0)
    assert np.max(np.abs(o["risk"] - g["risk"])) <= 1e-12
    assert np.array_equal(o["deviation"] < 0, g["deviation"] < 0)
    assert rel_err(o["deviation"], g["deviation"]) <= 1e-9
    assert np.max(np.abs(o["hazard"] - g["hazard"])) <= 1e-12
    assert o["traj"].shape == g["traj"].shape
    assert np.max(np.abs(o["traj"][:, :2] - g["traj"][:, :2])) <= 1e-3
    assert o["reconnecting_index"] == int(g["reconnecting_index"])
