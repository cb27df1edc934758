"""CoRa cost-ratio statistics: the host-side restatement in
planning-path_planning_b200/src/DyMuCoRa.hpp against the compiled, unmodified reference
(G.cpp:895-1038, H.hpp:110-394).  Pure host logic: runs without a GPU.

Without a device the B200 planner cannot build a cost map; computeCostMap reports failure but,
like the reference (G.cpp:151-153), has already recorded the look-up table, slope values and
locomotion modes, which is all CoRa needs."""
import numpy as np
import pytest

import scenarios


def _planners(pkg, ref_lib, nx=48, ny=40):
    syn = pkg.synthetic
    elev, terr = syn.mars_dem(ny, nx, seed=3)
    lut, slopes, locs = syn.default_lut()
    ref = scenarios.make_planner(ref_lib.DyMuPathPlanner, 1, nx, ny)
    assert ref.computeCostMap(lut, slopes, locs, elev, terr)
    plib = pkg.planner_lib()
    new = plib.DyMuPathPlanner(1.0, 1.5, 2.0, 1)
    new.nx, new.ny = nx, ny
    new.computeCostMap(lut, slopes, locs, elev, terr)   # table recorded even when this fails
    return ref, new, np.asarray(lut, dtype=np.float64)


def test_update_cost_matches_reference(pkg, ref_lib):
    ref, new, lut0 = _planners(pkg, ref_lib)
    taps_ref = scenarios.cora_feed(ref)
    taps_new = scenarios.cora_feed(new)
    assert len(taps_ref) == len(taps_new) == 6
    changed = False
    for (ratio_r, lut_r), (ratio_n, lut_n) in zip(taps_ref, taps_new):
        assert ratio_r.shape == ratio_n.shape and lut_r.shape == lut_n.shape == lut0.shape
        assert np.array_equal(ratio_r, ratio_n, equal_nan=True)
        assert np.array_equal(lut_r, lut_n, equal_nan=True)
        changed |= not np.array_equal(lut_r, lut0)
    assert changed, "the scenario never rewrote the table"
    assert len(taps_ref[-1][0]) == 3      # all four terrains traversed at the end
    assert len(taps_ref[1][0]) == 2       # terrain 2 still unknown: ratio chain skips it


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_update_cost_matches_reference_other_streams(pkg, ref_lib, seed):
    ref, new, _ = _planners(pkg, ref_lib)
    taps_ref = scenarios.cora_feed(ref, seed=seed, rounds=90, tap_every=15, weights=(3.0, 0.5))
    taps_new = scenarios.cora_feed(new, seed=seed, rounds=90, tap_every=15, weights=(3.0, 0.5))
    for (ratio_r, lut_r), (ratio_n, lut_n) in zip(taps_ref, taps_new):
        assert np.array_equal(ratio_r, ratio_n, equal_nan=True)
        assert np.array_equal(lut_r, lut_n, equal_nan=True)


def test_cora_argument_checks(pkg, ref_lib):
    ref, new, _ = _planners(pkg, ref_lib)
    for p in (ref, new):
        assert not p.initCoRaMethod(4, 2, [1.0])            # weights do not match the criteria
        assert p.initCoRaMethod(4, 2, [1.0, 1.0])
        assert not p.fillTerrainInfo(0, [1.0, 2.0, 3.0])    # wrong sample width
        assert p.fillTerrainInfo(0, [1.0, 2.0])
        assert len(p.computeCostRatio()) == 0               # nothing traversed yet
    assert np.array_equal(ref.updateCost(), new.updateCost())
    assert not new.fillTerrainInfo(9, [1.0, 2.0])           # reference: out-of-range access
