"""Generates the committed golden fixtures from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference to have been compiled by
`make -C oracle ref`):   python tests/golden/make_golden.py
The reference ships no golden vectors of its own; these are outputs
of the reference itself on seeded synthetic inputs, inputs included so that the fixtures do
not depend on the generator staying unchanged."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dymu_b200  # noqa: E402
import oracle  # noqa: E402
import scenarios as sc  # noqa: E402


def config1(ref, syn):
    """BASELINE.json configs[0]: 100x100, goal (75,75), start (20,25)."""
    n = 100
    elev, terr = syn.mars_dem(n, n, seed=1)
    lut, slopes, locs = syn.default_lut()
    p = sc.make_planner(ref.DyMuPathPlanner, 1, n, n)
    assert p.computeCostMap(lut, slopes, locs, elev, terr)
    ob = p.node_field(4)
    gi, gj = syn.free_interior_cell_near(ob, 75, 75)
    si, sj = syn.free_interior_cell_near(ob, 20, 25)
    assert p.setGoal(gi, gj)
    assert p.computeTotalCostMap(si, sj)
    out = dict(elevation=elev, terrain=terr, lut=lut, slopes=slopes, goal=np.array([gi, gj]),
               start=np.array([si, sj]), obstacle=ob.astype(np.uint8), slope=p.node_field(1),
               raw_cost=p.node_field(2), cost=p.node_field(3), total_cost_early=p.getTotalCostMatrix(),
               state_early=p.node_field(5).astype(np.uint8), path_early=p.getPath(si, sj))
    assert p.computeEntireTotalCostMap()
    out["total_cost_full"] = p.getTotalCostMatrix()
    out["path_full"] = p.getPath(si, sj)
    np.savez_compressed(os.path.join(HERE, "config1_100.npz"), **out)
    print("config1_100: %d closed cells, %d / %d waypoints" %
          (out["state_early"].sum(), len(out["path_early"]), len(out["path_full"])))


def repair(ref, syn, approach, name):
    n = 120
    elev, terr = syn.mars_dem(n, n, seed=3)
    lut, slopes, locs = syn.default_lut()
    p = sc.make_planner(ref.DyMuPathPlanner, approach, n, n)
    assert p.computeCostMap(lut, slopes, locs, elev, terr)
    ob = p.node_field(4)
    gi, gj = syn.free_interior_cell_near(ob, 96, 96)
    si, sj = syn.free_interior_cell_near(ob, 24, 24)
    assert p.setGoal(gi, gj) and p.computeEntireTotalCostMap()
    path = p.getPath(si, sj)
    r = sc.repair_scenario(p, syn, path)
    assert r["repaired"]
    np.savez_compressed(os.path.join(HERE, name), elevation=elev, terrain=terr, lut=lut,
                        slopes=slopes, goal=np.array([gi, gj]), start=np.array([si, sj]),
                        path=path, image=r["image"], centre=r["centre"], risk=r["risk"],
                        deviation=r["deviation"], traj=r["traj"], hazard=r["hazard"],
                        traff=r["traff"], reconnecting_index=np.array(r["reconnecting_index"]),
                        has_local=p.node_field(6).astype(np.uint8))
    print("%s: %d trajectory waypoints, %d risk cells" % (name, len(r["traj"]),
                                                          int((r["risk"] > 0).sum())))


if __name__ == "__main__":
    assert oracle.have_reference(), "build oracle/_ref first (make -C oracle ref)"
    ref = oracle.reference()
    syn = dymu_b200.load().synthetic
    config1(ref, syn)
    repair(ref, syn, 1, "repair_120_sweeping.npz")
    repair(ref, syn, 0, "repair_120_conservative.npz")
