"""Direct delivery of the total-cost matrix (dymu_set_total_cost_export): what it costs the solve
kernel and what a plan gains.  Run on a GPU box: python scripts/probe_export.py [n]"""
import importlib
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
pkg = importlib.import_module("planning-path_planning_b200")
api, syn = pkg.cuda_api, pkg.synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
elev, terr = syn.mars_dem(n, n, seed=7)
lut, slopes, locs = syn.default_lut()
dev = api.DeviceLayer(n, n, 1.0, 0.1)
dev.compute_cost_map(lut, slopes, len(locs), elev, terr)
ob = dev.download_plane_u8("obstacle")
goal = syn.free_interior_cell_near(ob, n // 2, n // 2)
cost_host = torch.empty((n, n), dtype=torch.float64).pin_memory()
dev.download_plane("cost", xform=api.XFORM_EFFECTIVE_COST, out=cost_host.numpy())
t_host = torch.empty((n, n), dtype=torch.float64).pin_memory()


def timed(fn, reps=6):
    fn()
    fn()
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        st = fn()
        torch.cuda.synchronize()
        out.append(((time.perf_counter() - t0) * 1e3, st))
    out.sort(key=lambda x: x[0])
    return out[len(out) // 2]


def resident():
    return dev.solve_total_cost([goal])


def resident_then_copy():
    st = dev.solve_total_cost([goal])
    dev.download_total_cost_begin(t_host.numpy(), xform=api.XFORM_INF_TO_MINUS1)
    dev.download_total_cost_end()
    return st


def streamed():
    st = dev.plan_streamed(cost_host.numpy(), goal)
    dev.download_total_cost_begin(t_host.numpy(), xform=api.XFORM_INF_TO_MINUS1)
    dev.download_total_cost_end()
    return st


for label, export in (("copy afterwards", False), ("direct delivery", True)):
    dev.set_total_cost_export(t_host.numpy() if export else None, xform=api.XFORM_INF_TO_MINUS1)
    for name, fn in (("resident solve only", resident), ("resident solve + matrix", resident_then_copy),
                     ("streamed plan + matrix", streamed)):
        ms, st = timed(fn)
        print("%-16s %-24s %7.2f ms wall, kernel %6.2f ms, %3d phases, tiles early %5d late %5d"
              % (label, name, ms, st["kernel_ms"], st["outer_iterations"], st["tiles_delivered_early"],
                 st["tiles_delivered_late"]), flush=True)
