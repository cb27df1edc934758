"""Developer probe: sweep of the priority-band width and the per-activation sweep cap."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dymu_b200
pkg = dymu_b200.load(); syn = pkg.synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
elev, terr = syn.mars_dem(n, n, seed=20261018)
lut, slopes, locs = syn.default_lut()
goal = None
grid = [(int(a), float(b), int(c)) for a, b, c in
        (t.split(":") for t in os.environ.get("PROBE_GRID", "").split(",") if t)] or \
       [(inner, band, 0) for inner in (32, 48, 64, 96) for band in (1.5, 2.0, 2.5, 3.0, 4.0, 6.0)]
for inner, band, budget in grid:
    if True:
        os.environ["DYMU_FIM_BAND"] = str(band)
        os.environ["DYMU_FIM_INNER"] = str(inner)
        os.environ["DYMU_FIM_BUDGET"] = str(budget)
        dev = pkg.cuda_api.DeviceLayer(n, n, 1.0, 0.1)
        dev.compute_cost_map(lut, slopes, len(locs), elev, terr)
        if goal is None:
            ob = dev.download_plane_u8("obstacle")
            goal = syn.free_interior_cell_near(ob, n // 2, n // 2)
        best = None
        for rep in range(3):
            st = dev.solve_total_cost([goal])
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        print("inner %3d band %.1f budget %3d: %.2f ms, %d phases, %d activations, %.1f sweeps/activation" %
              (inner, band, budget, best["kernel_ms"], best["outer_iterations"], best["tile_activations"],
               best["inner_iterations"] / max(best["tile_activations"], 1)), flush=True)
        del dev
