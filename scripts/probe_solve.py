"""Developer probe: time the tile-FIM total-cost solve on a synthetic map (not the bench)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dymu_b200

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4096)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--kind", default="mars")
ap.add_argument("--check", action="store_true")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--nopath", action="store_true")
args = ap.parse_args()

pkg = dymu_b200.load()
syn = pkg.synthetic
n = args.n
t0 = time.time()
dev = pkg.cuda_api.DeviceLayer(n, n, 1.0, 0.1)
if args.kind == "mars":
    elev, terr = syn.mars_dem(n, n)
    lut, slopes, locs = syn.default_lut()
    t1 = time.time()
    dev.compute_cost_map(lut, slopes, len(locs), elev, terr)
else:
    cost = syn.smooth_cost_map(n, n)
    t1 = time.time()
    dev.set_cost_map(cost)
dev.synchronize()
t2 = time.time()
ob = dev.download_plane_u8("obstacle")
print("gen %.2fs costmap(+H2D) %.3fs obstacle frac %.4f tile/pitch/rows %s" %
      (t1 - t0, t2 - t1, ob.mean(), dev.geometry()), flush=True)
rng = np.random.default_rng(11)
goals = [syn.free_interior_cell_near(ob, n // 2, n // 2)]
while len(goals) < args.batch:
    goals.append(syn.free_interior_cell_near(ob, int(rng.uniform(0.1, 0.9) * n),
                                             int(rng.uniform(0.1, 0.9) * n)))
if args.batch > 1:
    dev.reserve_slots(args.batch)
for r in range(args.reps):
    st = dev.solve_total_cost(goals)
    reached = dev.count_reached()
    print("rep %d: kernel %.3f ms reset %.3f ms outer %d tiles %d deferred %d inner/tile %.1f updates %.3e (%.1f/cell) reached %.4f conv %d"
          % (r, st["kernel_ms"], st["reset_ms"], st["outer_iterations"], st["tile_activations"],
             st["tiles_deferred"], st["inner_iterations"] / max(1, st["tile_activations"]),
             st["cell_updates"], st["cell_updates"] / (n * n * len(goals)), reached / (n * n),
             st["converged"]), flush=True)
if args.nopath:
    sys.exit(0)
si, sj = syn.free_interior_cell_near(ob, n // 8, n // 8)
t = time.time()
wps, status = dev.extract_global_path(float(si), float(sj), 0.4, goals[0][0], goals[0][1])
print("path: %d wps status %d in %.3f ms" % (len(wps), status, (time.time() - t) * 1e3))
if args.check:
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle
    po = oracle.Port(1.0, 1.5, 2.0, 1)
    po.initGlobalLayer(1.0, 0.1, n, n)
    if args.kind == "mars":
        po.computeCostMap(lut, slopes, locs, elev, terr)
    else:
        po.setCostMap(cost)
    po.setGoal(*goals[0])
    t = time.time()
    po.computeEntireTotalCostMap(heap=True)
    print("heap oracle %.2fs" % (time.time() - t))
    T, To = dev.download_total_cost(), po.plane("total_cost")
    fin = np.isfinite(To) & (To > 0)
    print("mask equal", np.array_equal(np.isinf(T), np.isinf(To)), "max rel err",
          np.max(np.abs(T[fin] - To[fin]) / To[fin]), "bit-equal frac", (T == To).mean())
