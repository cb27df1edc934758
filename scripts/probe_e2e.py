"""Developer probe: per-call wall time of the class end-to-end step (bench.py's e2e leg)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dymu_b200
pkg = dymu_b200.load(); syn = pkg.synthetic
n = 4096
elev, terr = syn.mars_dem(n, n)
lut, slopes, locs = syn.default_lut()
dev = pkg.cuda_api.DeviceLayer(n, n, 1.0, 0.1)
dev.compute_cost_map(lut, slopes, len(locs), elev, terr)
ob = dev.download_plane_u8("obstacle")
goal = syn.free_interior_cell_near(ob, n // 2, n // 2)
start = syn.free_interior_cell_near(ob, n // 8, n // 8)
cost_host = torch.empty((n, n), dtype=torch.float64).pin_memory()
dev.download_plane("cost", xform=pkg.cuda_api.XFORM_EFFECTIVE_COST, out=cost_host.numpy())
t_host = torch.empty((n, n), dtype=torch.float64).pin_memory()
dev.close()
pl = pkg.DyMuPathPlanner(1.0, 1.5, 2.0, 1)
assert pl.initGlobalLayer(1.0, 0.1, n, n)
assert pl.setCostMapFlat(cost_host.numpy()) and pl.setGoal(float(goal[0]), float(goal[1]))
assert pl.setTotalCostMatrixTarget(t_host.numpy())
for it in range(6):
    t = [time.perf_counter()]
    pl.setCostMapFlat(cost_host.numpy()); t.append(time.perf_counter())
    pl.computeEntireTotalCostMap(); t.append(time.perf_counter())
    p = pl.getPath(float(start[0]), float(start[1])); t.append(time.perf_counter())
    pl.getTotalCostMatrixFlat(t_host.numpy()); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print("step %d: setCostMap %.3f  computeEntire %.3f  getPath %.3f (in-library %.3f)  getMatrix %.3f  total %.3f ms"
          % (it, d[0], d[1], d[2], pl.last_call_seconds * 1e3, d[3], d.sum()), flush=True)
