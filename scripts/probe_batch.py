"""Developer probe: batched goal queries on one shared map (BASELINE configs[3] shape) and
device-side timing of the path kernel."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dymu_b200
ap = argparse.ArgumentParser()
ap.add_argument("--size", dest="n", type=int, default=2048)
ap.add_argument("--batches", default="1,4,16,64")
a = ap.parse_args()
pkg = dymu_b200.load(); syn = pkg.synthetic
n = a.n
dev = pkg.cuda_api.DeviceLayer(n, n)
elev, terr = syn.mars_dem(n, n)
lut, slopes, locs = syn.default_lut()
dev.compute_cost_map(lut, slopes, len(locs), elev, terr)
ob = dev.download_plane_u8("obstacle")
rng = np.random.default_rng(11)
free = np.argwhere(~syn.free_interior_cell_near.__globals__['np'].zeros(1, bool)) if False else None
def rand_goal():
    return syn.free_interior_cell_near(ob, int(rng.uniform(0.05, 0.95) * n), int(rng.uniform(0.05, 0.95) * n))
for B in [int(x) for x in a.batches.split(",")]:
    dev.reserve_slots(B)
    goals = [rand_goal() for _ in range(B)]
    dev.solve_total_cost(goals)
    best = None
    for rep in range(3):
        st = dev.solve_total_cost(goals)
        best = st if best is None or st["kernel_ms"] < best["kernel_ms"] else best
    print("batch %3d: kernel %.2f ms -> %.1f solves/s, %.2f ms/solve, updates/cell %.1f, outer %d" % (
        B, best["kernel_ms"], B / best["kernel_ms"] * 1e3, best["kernel_ms"] / B,
        best["cell_updates"] / (n * n * B), best["outer_iterations"]), flush=True)
# path kernel device time
dev.reserve_slots(1)
g = syn.free_interior_cell_near(ob, n // 2, n // 2)
dev.solve_total_cost([g])
s = syn.free_interior_cell_near(ob, n // 8, n // 8)
for rep in range(3):
    dev.event_record(0)
    wps, status = dev.extract_global_path(float(s[0]), float(s[1]), 0.4, g[0], g[1], cap=1 << 15)
    dev.event_record(1)
    print("path: %d wps status %d device+copy %.3f ms" % (len(wps), status, dev.event_elapsed_ms(0, 1)))
