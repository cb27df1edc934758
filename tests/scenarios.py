"""Shared scenario drivers: the same call sequence is issued to any object exposing the
DyMuPathPlanner method names (reference .so, C port, B200 library)."""
import numpy as np


def make_planner(factory, approach, nx, ny, offset=(0.0, 0.0), risk_distance=1.0,
                 reconnect_distance=1.5, risk_ratio=2.0):
    p = factory(risk_distance, reconnect_distance, risk_ratio, approach)
    assert p.initGlobalLayer(1.0, 0.1, nx, ny, offset)
    return p


def obstacle_plane(p):
    if hasattr(p, "node_field"):
        return p.node_field(4)
    return p.plane("isObstacle").astype(np.float64)


def global_scenario(p, syn, nx, ny, seed, goal_frac=(0.8, 0.8), start_frac=(0.2, 0.2),
                    entire=False, use_cost_map=False, offset=(0.0, 0.0)):
    """SURVEY.md section 8d config 1/2 style: cost map -> goal -> total cost -> path."""
    if use_cost_map:
        assert p.setCostMap(syn.smooth_cost_map(ny, nx, seed=seed))
    else:
        elev, terr = syn.mars_dem(ny, nx, seed=seed)
        lut, slopes, locs = syn.default_lut()
        assert p.computeCostMap(lut, slopes, locs, elev, terr)
    ob = obstacle_plane(p)
    gi, gj = syn.free_interior_cell_near(ob, int(nx * goal_frac[0]), int(ny * goal_frac[1]))
    si, sj = syn.free_interior_cell_near(ob, int(nx * start_frac[0]), int(ny * start_frac[1]))
    gx, gy, sx, sy = gi + offset[0], gj + offset[1], si + offset[0], sj + offset[1]
    assert p.setGoal(gx, gy)
    ok = p.computeEntireTotalCostMap() if entire else p.computeTotalCostMap(sx, sy)
    out = dict(ok=ok, goal=(gi, gj), start=(si, sj), obstacle=ob)
    out["T"] = p.getTotalCostMatrix()
    out["cost"] = p.getGlobalCostMatrix()
    out["path"] = p.getPath(sx, sy)
    return out


def repair_scenario(p, syn, path, seed=7, frame=120, res=0.1, disc_wp=12, disc_radius=0.8):
    """Config 2: a traversability frame centred on path[0] with an obstacle disc on the path
    plus three random discs; then computeLocalPlanning."""
    c = path[0, :2]
    d = path[min(disc_wp, len(path) - 2), :2]
    rng = np.random.default_rng(seed)
    discs = [(d[0], d[1], disc_radius)]
    discs += [(c[0] + rng.uniform(-4, 4), c[1] + rng.uniform(-4, 4), rng.uniform(0.2, 0.5))
              for _ in range(3)]
    img = syn.obstacle_frame(frame, frame, res, c, discs)
    repaired, traj, local_time = p.computeLocalPlanning(c[0], c[1], img, res)
    return dict(repaired=repaired, traj=traj, local_time=local_time, centre=c, image=img,
                risk=p.getRiskMatrix(c[0], c[1]), deviation=p.getDeviationMatrix(c[0], c[1]),
                hazard=p.getHazardDensityMatrix(), traff=p.getTrafficabilityMatrix(),
                reconnecting_index=p.getReconnectingIndex())


def cora_feed(p, n_terrains=4, n_criteria=2, weights=(1.0, 2.0), seed=5, rounds=60, tap_every=10):
    """CoRa (G.cpp:895-1038): stream seeded traverse samples terrain by terrain and tap
    updateCost()/computeCostRatio() along the way.  The distributions drift and jump so that the
    accept (Student / Cochran), reject, and rejected-takes-over branches of dataAnalysis
    (H.hpp:234-309) all run.  initCoRaMethod needs the table of a previous computeCostMap."""
    assert p.initCoRaMethod(n_terrains, n_criteria, list(weights))
    rng = np.random.default_rng(seed)
    taps = []
    for r in range(rounds):
        for t in range(n_terrains):
            if t == 2 and r < rounds // 2:
                continue                      # terrain 2 is met late: ratios must skip over it
            base = np.array([0.2 + 0.15 * t, 30.0 + 12.0 * t])
            spread = np.array([0.02, 2.0]) * (1.0 + 0.5 * t)
            if t == 1 and r >= rounds // 3:
                # lasting drop (the Student test is one-sided: only lower means are rejected)
                # -> rejected info piles up and eventually replaces the kept statistics
                base, spread = base * 0.6, spread * 0.6
            if t == 0 and r >= rounds // 2:
                base = base * 1.5             # lasting rise: accepted, drags the mean along
            if t == 3 and r % 7 == 3:
                spread = spread * 6.0         # noisy stretch -> Cochran branch
            for _ in range(int(rng.integers(2, 6))):
                sample = base + spread * rng.standard_normal(n_criteria)
                if rng.random() < 0.15:
                    sample[int(rng.integers(0, n_criteria))] = -1.0   # "no reading"
                assert p.fillTerrainInfo(t, sample)
        if r % tap_every == tap_every - 1:
            taps.append((p.computeCostRatio(), p.updateCost()))
    return taps
