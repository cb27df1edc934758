"""Pins the plain-C restatement (oracle/dymu_oracle.c) against the UNMODIFIED reference
compiled into oracle/_ref (the reference ships no golden vectors, so
the reference itself is the pin).  Everything here is CPU-only."""
import numpy as np
import pytest

import scenarios as sc


@pytest.mark.parametrize("ny,nx,seed,entire", [(100, 100, 1, False), (129, 257, 2, True),
                                               (60, 90, 3, False), (200, 200, 4, True)])
def test_port_global_layer_bit_exact(pkg, oracle_mod, ref_lib, ny, nx, seed, entire):
    syn = pkg.synthetic
    a = sc.global_scenario(sc.make_planner(ref_lib.DyMuPathPlanner, 1, nx, ny), syn, nx, ny, seed,
                           entire=entire)
    b = sc.global_scenario(sc.make_planner(oracle_mod.Port, 1, nx, ny), syn, nx, ny, seed,
                           entire=entire)
    assert a["ok"] == b["ok"]
    assert np.array_equal(a["obstacle"], b["obstacle"])
    assert np.array_equal(a["T"], b["T"])
    assert np.array_equal(a["cost"], b["cost"])
    assert np.array_equal(a["path"], b["path"])


def test_port_with_offset_and_cost_map(pkg, oracle_mod, ref_lib):
    syn, off = pkg.synthetic, (12.0, -7.0)
    a = sc.global_scenario(sc.make_planner(ref_lib.DyMuPathPlanner, 1, 120, 90, off), syn, 120, 90,
                           5, use_cost_map=True, offset=off)
    b = sc.global_scenario(sc.make_planner(oracle_mod.Port, 1, 120, 90, off), syn, 120, 90, 5,
                           use_cost_map=True, offset=off)
    assert a["ok"] and b["ok"]
    assert np.array_equal(a["T"], b["T"])
    # z is uninitialised memory in the reference when no elevation was ever set
    assert np.array_equal(a["path"][:, [0, 1, 3]], b["path"][:, [0, 1, 3]])


@pytest.mark.parametrize("approach", [0, 1])
def test_port_local_repair_bit_exact(pkg, oracle_mod, ref_lib, approach):
    syn, n = pkg.synthetic, 200
    res = []
    for factory in (ref_lib.DyMuPathPlanner, oracle_mod.Port):
        p = sc.make_planner(factory, approach, n, n)
        g = sc.global_scenario(p, syn, n, n, seed=3)
        r = sc.repair_scenario(p, syn, g["path"])
        res.append((g, r))
    (ga, ra), (gb, rb) = res
    assert ra["repaired"] and rb["repaired"]
    for k in ("traj", "risk", "deviation", "hazard", "traff"):
        assert np.array_equal(ra[k], rb[k]), k
    assert ra["reconnecting_index"] == rb["reconnecting_index"]


def test_heap_variant_agrees_with_literal(pkg, oracle_mod):
    syn, n = pkg.synthetic, 300
    elev, terr = syn.mars_dem(n, n, seed=8)
    lut, slopes, locs = syn.default_lut()
    p = oracle_mod.Port(1.0, 1.5, 2.0, 1)
    p.initGlobalLayer(1.0, 0.1, n, n)
    p.computeCostMap(lut, slopes, locs, elev, terr)
    gi, gj = syn.free_interior_cell_near(p.plane("isObstacle"), n // 2, n // 2)
    p.setGoal(gi, gj)
    p.computeEntireTotalCostMap()
    T1 = p.plane("total_cost")
    p.computeEntireTotalCostMap(heap=True)
    T2 = p.plane("total_cost")
    assert np.array_equal(np.isinf(T1), np.isinf(T2))
    fin = np.isfinite(T1) & (T1 > 0)
    assert np.max(np.abs(T1[fin] - T2[fin]) / T1[fin]) <= 1e-14
    assert p.fixed_point_violations(T1, 1e-12) == 0


def test_analytic_uniform_cost(oracle_mod):
    """Known answer: uniform cost c => along the goal's row/column T = distance * c exactly."""
    n, c = 65, 1.5
    p = oracle_mod.Port(1.0, 1.5, 2.0, 1)
    p.initGlobalLayer(1.0, 0.1, n, n)
    p.setCostMap(np.full((n, n), c))
    p.setGoal(32, 32)
    p.computeEntireTotalCostMap()
    T = p.plane("total_cost")
    assert np.array_equal(T[32, 32:], c * np.arange(0, n - 32))
    assert np.array_equal(T[:33, 32][::-1], c * np.arange(0, 33))
    assert np.all(np.isfinite(T))
