"""Developer probe: config 2 (1000x1000 + local repair) timings, reference vs B200 drop-in.
DYMU_TIMING=1 prints the per-stage wall times of the drop-in's local layer."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dymu_b200, oracle, scenarios as sc
pkg = dymu_b200.load(); syn = pkg.synthetic
n = 1000
labels = [("b200", pkg.DyMuPathPlanner)]
if "--ref" in sys.argv:
    labels.insert(0, ("reference", oracle.reference().DyMuPathPlanner))
for approach, name in ((1, "SWEEPING"), (0, "CONSERVATIVE")):
    for label, factory in labels:
        for rep in range(3 if label == "b200" else 1):
            p = sc.make_planner(factory, approach, n, n)
            t0 = time.perf_counter()
            g = sc.global_scenario(p, syn, n, n, seed=20261018)
            t1 = time.perf_counter()
            sys.stderr.write("---- %s %s rep %d\n" % (name, label, rep)); sys.stderr.flush()
            t2 = time.perf_counter()
            r = sc.repair_scenario(p, syn, g["path"], disc_wp=10)
            print("%-12s %-9s global scenario %.3f s | repair scenario (computeLocalPlanning + taps) %.2f ms | in-library %.2f ms (repaired=%s, %d wps)"
                  % (name, label, t1 - t0, (time.perf_counter() - t2) * 1e3, r["local_time"] * 1e3, r["repaired"], len(r["traj"])), flush=True)
            p.close()
