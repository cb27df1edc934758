"""Developer probe: config 2 (1000x1000 + local repair) timings, reference vs B200 drop-in."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dymu_b200, oracle, scenarios as sc
pkg = dymu_b200.load(); syn = pkg.synthetic
n = 1000
for approach, name in ((1, "SWEEPING"), (0, "CONSERVATIVE")):
    for label, factory in (("reference", oracle.reference().DyMuPathPlanner), ("b200", pkg.DyMuPathPlanner)):
        p = sc.make_planner(factory, approach, n, n)
        t0 = time.perf_counter()
        g = sc.global_scenario(p, syn, n, n, seed=20261018)
        t1 = time.perf_counter()
        r = sc.repair_scenario(p, syn, g["path"], disc_wp=10)
        t_rep = p.last_call_seconds if hasattr(p, "last_call_seconds") else float("nan")
        # second identical call: nothing new -> no repair, measures the ingest-only path
        t2 = time.perf_counter()
        print("%-12s %-9s global scenario %.3f s | computeLocalPlanning %.2f ms (repaired=%s, %d wps, internal localTime %.2f ms)"
              % (name, label, t1 - t0, r["local_time"] * 1e3 if False else (time.perf_counter() - t2) * 0 + r["local_time"] * 1e3,
                 r["repaired"], len(r["traj"]), r["local_time"] * 1e3), flush=True)
