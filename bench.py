#!/usr/bin/env python
"""bench.py -- DyMu total-cost propagation on B200: one JSON line per run.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on and the largest
that is meant for one GPU): a 4096x4096 synthetic Mars-like DEM (seeded fBm + craters),
cost map built by the device cost-map pipeline (computeCostMap), one goal at the free
interior cell nearest the centre.

A "step" = one plan = full total-cost-map solve (computeEntireTotalCostMap equivalent:
reset + tile-FIM kernel) + gradient-descent path extraction from a fixed start.

  value  : plans/s with the cost planes already resident in HBM (device time, CUDA events
           on the library's stream, max over ranks).  ms_per_step is BASELINE's "ms per
           4096^2 total-cost-map solve" (+ path).
  e2e    : the same through the C ABI with HOST buffers inside the timed region: H2D of the
           cost plane (setCostMap), solve, path, D2H of the total-cost matrix
           (getTotalCostMatrix) every step.
  N > 1  : one process per GPU (torchrun); every rank plans on its own copy of the map
           with its own goal -- independent queries, no data-path collective ("weak").

--impl reference times the UNMODIFIED reference (oracle/_ref) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "plans_per_s"
UNIT = "plans/s"
CELL_SWEEP_BYTES = 24   # per cell of an activated tile: read T, read C_eff, write T (fp64)
SOLVE_BYTES_PER_CELL = 16  # lower bound per solve: read C_eff once, write T once


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="plan4096", choices=["plan4096", "queries2048"],
                    help="plan4096 = BASELINE configs[2] (default, the headline); queries2048 = "
                         "configs[3]: batches of independent goal queries on one shared 2048^2 map, "
                         "sharded across the ranks")
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seed", type=int, default=20261018)
    return ap.parse_args()


def dist_env():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_workload(pkg, n, seed):
    syn = pkg.synthetic
    elev, terr = syn.mars_dem(n, n, seed=seed)
    lut, slopes, locs = syn.default_lut()
    return elev, terr, lut, slopes, locs


# ---------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dymu_b200
    import oracle
    pkg = dymu_b200.load()
    n = args.size
    elev, terr, lut, slopes, locs = build_workload(pkg, n, args.seed)
    # bounded sample: the reference's vector narrow band scales ~x16 per doubling
    # (12.6 s @1000^2, 3152 s @4096^2), so each step solves the central crop
    # whose single-thread time fits the budget
    budget_s = max(2.0, 150.0 / max(1, args.steps + args.warmup))
    crop = int(1000.0 * (budget_s / 14.0) ** 0.25)
    crop = max(256, min(1024, crop, n))
    lo = (n - crop) // 2
    e_c = np.ascontiguousarray(elev[lo:lo + crop, lo:lo + crop])
    t_c = np.ascontiguousarray(terr[lo:lo + crop, lo:lo + crop])
    use_ref = oracle.have_reference()
    cores = min(os.cpu_count() or 1, 32)
    if use_ref:
        lib = oracle.reference()
        make = lambda: lib.DyMuPathPlanner(1.0, 1.5, 2.0, 1)
        kind = "reference"
    else:
        oracle.build(ref=False, port=True)
        make = lambda: oracle.Port(1.0, 1.5, 2.0, 1)
        kind = "port"
    planners, goals = [], []
    rng = np.random.default_rng(args.seed + 1)
    for c in range(cores):
        p = make()
        p.initGlobalLayer(1.0, 0.1, crop, crop)
        p.computeCostMap(lut, slopes, locs, e_c, t_c)
        ob = p.node_field(4) if hasattr(p, "node_field") else p.plane("isObstacle")
        gi, gj = pkg.synthetic.free_interior_cell_near(
            ob, crop // 2 + int(rng.integers(-crop // 8, crop // 8 + 1)),
            crop // 2 + int(rng.integers(-crop // 8, crop // 8 + 1)))
        si, sj = pkg.synthetic.free_interior_cell_near(ob, crop // 8, crop // 8)
        assert p.setGoal(gi, gj)
        planners.append((p, si, sj))

    def one(p, si, sj):
        p.computeEntireTotalCostMap()
        p.getPath(float(si), float(sj))

    def step():
        ts = [threading.Thread(target=one, args=t) for t in planners]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ms_per_step = dt / args.steps * 1e3
    frac = (crop * crop) / float(n * n)
    value = cores * frac / (dt / args.steps)
    sample = ("%d threads, each one full solve + path on the central %dx%d crop (%.4f of the "
              "%dx%d cells) per step; value = crop-fraction plans/s, i.e. assumes time linear in "
              "cells, which flatters the reference (its narrow-band scan scales ~x16 per "
              "doubling: 3152 s measured for one 4096^2 solve in BASELINE.md)"
              % (cores, crop, crop, frac, n, n))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "configs[2]: %dx%d synthetic Mars-like DEM cost map, FMM total-cost "
                               "solve + path, 1 goal" % (n, n), "seed": args.seed,
                   "reference_sample_crop": crop},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def cpu_baseline_leg(pkg, elev, terr, lut, slopes, locs, n):
    """Reference (oracle/_ref) or port, 1 thread, bounded sample (~15 s): central 1024^2 crop."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    crop = min(1024, n)
    lo = (n - crop) // 2
    e_c = np.ascontiguousarray(elev[lo:lo + crop, lo:lo + crop])
    t_c = np.ascontiguousarray(terr[lo:lo + crop, lo:lo + crop])
    if oracle.have_reference():
        p = oracle.reference().DyMuPathPlanner(1.0, 1.5, 2.0, 1)
        kind = "reference"
    else:
        oracle.build(ref=False, port=True)
        p = oracle.Port(1.0, 1.5, 2.0, 1)
        kind = "port"
    p.initGlobalLayer(1.0, 0.1, crop, crop)
    p.computeCostMap(lut, slopes, locs, e_c, t_c)
    ob = p.node_field(4) if hasattr(p, "node_field") else p.plane("isObstacle")
    gi, gj = pkg.synthetic.free_interior_cell_near(ob, crop // 2, crop // 2)
    p.setGoal(gi, gj)
    t0 = time.perf_counter()
    p.computeEntireTotalCostMap()
    p.getPath(float(crop // 8), float(crop // 8))
    dt = time.perf_counter() - t0
    frac = (crop * crop) / float(n * n)
    # the C port with a heap (not the reference's algorithmic complexity) on the full map
    fast = None
    try:
        q = oracle.Port(1.0, 1.5, 2.0, 1)
        q.initGlobalLayer(1.0, 0.1, n, n)
        q.computeCostMap(lut, slopes, locs, elev, terr)
        g2 = pkg.synthetic.free_interior_cell_near(q.plane("isObstacle"), n // 2, n // 2)
        q.setGoal(*g2)
        t1 = time.perf_counter()
        q.computeEntireTotalCostMap(heap=True)
        fast = time.perf_counter() - t1
        q.close()
    except Exception:
        fast = None
    return {
        "value": frac / dt, "unit": UNIT, "cores": 1, "kind": kind,
        "sample": ("1 thread, one full solve + path on the central %dx%d crop (%.4f of the cells) "
                   "in %.2f s; value = crop-fraction plans/s (linear-in-cells extrapolation, "
                   "optimistic for the reference)" % (crop, crop, frac, dt)),
        "sample_seconds": dt,
        "heap_port_full_map_seconds": fast,
    }


# ---------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------
def _measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one k_fim launch on this workload, from the
    committed `ncu --set full` capture (profiles/); None when the summary is not there."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles",
                               "r1_k_fim_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def run_b200(args):
    import torch
    import dymu_b200
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    pkg = dymu_b200.load()
    n = args.size
    elev, terr, lut, slopes, locs = build_workload(pkg, n, args.seed)
    dev = pkg.cuda_api.DeviceLayer(n, n, 1.0, 0.1, device=local_rank)
    dev.compute_cost_map(lut, slopes, len(locs), elev, terr)
    ob = dev.download_plane_u8("obstacle")
    syn = pkg.synthetic
    rng = np.random.default_rng(args.seed + 17 * rank)
    if rank == 0:
        goal = syn.free_interior_cell_near(ob, n // 2, n // 2)
    else:  # independent query per rank
        goal = syn.free_interior_cell_near(ob, int(rng.uniform(0.3, 0.7) * n),
                                           int(rng.uniform(0.3, 0.7) * n))
    start = syn.free_interior_cell_near(ob, n // 8, n // 8)
    tile, pitch, rows = dev.geometry()

    def plan():
        st = dev.solve_total_cost([goal])
        wps, status = dev.extract_global_path(float(start[0]), float(start[1]), 0.4, goal[0],
                                              goal[1])
        return st, len(wps), status

    # ---- resident-input arm -----------------------------------------------------------
    for _ in range(args.warmup):
        st, nwp, status = plan()
    reached = dev.count_reached()
    assert reached >= 0.90 * n * n, "goal is walled in: only %.3f of the map reached" % (
        reached / float(n * n))

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = dev.launches
    kernel_ms, tiles, updates, outers = [], [], [], []
    dev.event_record(0)
    for _ in range(args.steps):
        st, nwp, status = plan()
        kernel_ms.append(st["kernel_ms"])
        tiles.append(st["tile_activations"])
        updates.append(st["cell_updates"])
        outers.append(st["outer_iterations"])
    dev.event_record(1)
    total_ms = dev.event_elapsed_ms(0, 1)
    launches = dev.launches - launches0
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end arm: host buffers in, host matrix out, every step --------------------
    cost_host = torch.empty((n, n), dtype=torch.float64).pin_memory()
    dev.download_plane("cost", out=cost_host.numpy())
    t_host = torch.empty((n, n), dtype=torch.float64).pin_memory()

    def plan_e2e():
        # setCostMap + computeEntireTotalCostMap from the pinned host buffer; the upload of the
        # rows away from the goal runs on the copy stream behind the first solver phases
        dev.plan_streamed(cost_host.numpy(), goal)
        # getTotalCostMatrix read-back runs on the copy stream while the path is extracted
        dev.download_total_cost_begin(t_host.numpy(), xform=pkg.cuda_api.XFORM_INF_TO_MINUS1)
        dev.extract_global_path(float(start[0]), float(start[1]), 0.4, goal[0], goal[1])
        dev.download_total_cost_end()

    for _ in range(min(2, args.warmup)):
        plan_e2e()
    barrier()
    dev.event_record(2)
    for _ in range(args.steps):
        plan_e2e()
    dev.event_record(3)
    e2e_ms = dev.event_elapsed_ms(2, 3)
    barrier()

    if distributed:
        t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    ms_per_step = total_ms / args.steps
    value = world * 1e3 / ms_per_step
    e2e_value = world * 1e3 / (e2e_ms / args.steps)
    peak, peak_src = measured_peak()
    k_ms = statistics.mean(kernel_ms)
    # (i) 24 B per cell update (read C_eff, read T, write T) x the cell
    # updates one launch performs -- the figure `achieved` is built from; (ii) 16 B per reached
    # cell per solve, the lower bound.  The tile-level figure counts only what an activation has
    # to move between HBM/L2 and shared memory.
    algo_bytes = statistics.mean(updates) * CELL_SWEEP_BYTES
    achieved = algo_bytes / (k_ms * 1e-3) / 1e9
    tile_bytes = statistics.mean(tiles) * tile * tile * CELL_SWEEP_BYTES
    solve_bytes = SOLVE_BYTES_PER_CELL * reached
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": "configs[2]: %dx%d synthetic Mars-like DEM cost map (device computeCostMap), "
                        "isotropic upwind FMM total-cost solve + gradient-descent path, 1 goal per "
                        "GPU" % (n, n),
            "seed": args.seed, "tile": tile, "obstacle_fraction": float(ob.mean()),
            "reached_fraction": reached / float(n * n), "waypoints": nwp,
            "l2": "inputs larger than L2 (each fp64 plane %.0f MB > 126 MB L2)" % (n * n * 8 / 1e6),
            "parallelism": "1 independent plan per GPU, no collective",
        },
        "solve_kernel_ms": k_ms,
        "gcell_updates_per_s": statistics.mean(updates) / (k_ms * 1e-3) / 1e9,
        "updates_per_cell": statistics.mean(updates) / float(n * n),
        "outer_iterations": statistics.mean(outers),
        "roofline": {
            "bound": "hbm", "kernel": "k_fim<%d,0> (tile FIM sweep)" % tile,
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": _measured_traffic(), "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algo_bytes,
            "definition": "SURVEY 8(d)(i): cell updates x 24 B (read C_eff, read T, write T) / kernel "
                          "time.  The updates of an activated tile run in shared memory, so the DRAM "
                          "bytes actually moved (`traffic`, ncu) are ~1 %% of this: the kernel is "
                          "bounded by dependency latency and fp64 issue, not by HBM (DESIGN.md "
                          "section 5)",
            "tile_level": {"bytes": tile_bytes, "achieved": tile_bytes / (k_ms * 1e-3) / 1e9,
                           "frac": tile_bytes / (k_ms * 1e-3) / 1e9 / peak,
                           "definition": "tile activations x %d^2 cells x 24 B: what has to cross "
                                         "between L2/HBM and shared memory" % tile},
            "solve_level": {"bytes": solve_bytes,
                            "achieved": solve_bytes / (k_ms * 1e-3) / 1e9,
                            "frac": solve_bytes / (k_ms * 1e-3) / 1e9 / peak,
                            "definition": "SURVEY 8(d)(ii): 16 B x reached cells / solve time "
                                          "(lower bound: 41 us at peak)"},
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": n * n * 8, "d2h_bytes_per_step": n * n * 8 + nwp * 40},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    # the streaming (HBM-bound) stencil kernels of the same path, timed once each
    try:
        sm = dev.time_stencils()
        cells, pcells = float(n * n), float(pitch * rows)
        spec = [("k_fill_f64 (total-cost reset)", sm[0], 8 * cells),
                ("k_ceff (C_eff build)", sm[1], 33 * pcells),
                ("k_readback (inf -> -1)", sm[2], 16 * cells),
                ("k_set_cost_map (obstacle mask)", sm[3], 8 * cells)]
        line["stencils"] = {nm: {"ms": ms_, "algorithmic_bytes": b, "gbs": b / (ms_ * 1e-3) / 1e9,
                                 "frac_of_measured_peak": b / (ms_ * 1e-3) / 1e9 / peak}
                            for nm, ms_, b in spec if ms_ > 0}
    except Exception as e:  # timing helper is informational
        line["stencils"] = {"error": str(e)}
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline_leg(pkg, elev, terr, lut, slopes, locs, n)
    print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


def run_queries(args):
    """configs[3]: `--queries` independent goal queries (full solve + path each) on one shared
    2048x2048 map, block-sharded over the ranks (no data-path collective), `--batch` goals per
    launch.  One JSON line; a step = this rank's whole share."""
    import torch
    import dymu_b200
    rank, local_rank, world = dist_env()
    torch.cuda.set_device(local_rank)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    pkg = dymu_b200.load()
    syn, sh = pkg.synthetic, pkg.sharding
    n = 2048
    elev, terr, lut, slopes, locs = build_workload(pkg, n, args.seed)
    dev = pkg.cuda_api.DeviceLayer(n, n, 1.0, 0.1, device=local_rank)
    dev.compute_cost_map(lut, slopes, len(locs), elev, terr)
    ob = dev.download_plane_u8("obstacle")
    rng_g, rng_s = np.random.default_rng(11), np.random.default_rng(13)
    goals = [syn.free_interior_cell_near(ob, int(rng_g.uniform(0.03, 0.97) * n),
                                         int(rng_g.uniform(0.03, 0.97) * n)) for _ in range(args.queries)]
    starts = [syn.free_interior_cell_near(ob, int(rng_s.uniform(0.03, 0.97) * n),
                                          int(rng_s.uniform(0.03, 0.97) * n)) for _ in range(args.queries)]
    mine = list(sh.shard_queries(args.queries, world, rank))
    B = max(1, min(args.batch, len(mine)))
    dev.reserve_slots(B)

    def share():
        nwp = 0
        for lo in range(0, len(mine), B):
            idx = mine[lo:lo + B]
            dev.solve_total_cost([goals[q] for q in idx])
            paths, _ = dev.extract_global_path_batch(
                list(range(len(idx))), [[float(starts[q][0]), float(starts[q][1])] for q in idx], 0.4,
                [goals[q] for q in idx], cap=1 << 14)
            nwp += sum(len(w) for w in paths)
        return nwp

    for _ in range(max(1, min(args.warmup, 1))):
        share()
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    launches0 = dev.launches
    dev.event_record(0)
    for _ in range(args.steps):
        nwp = share()
    dev.event_record(1)
    ms = dev.event_elapsed_ms(0, 1)
    launches = dev.launches - launches0
    if distributed:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": args.queries * args.steps / (ms * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "configs[3]: %d independent goal queries (solve + path) on one shared "
                                   "2048x2048 map, %d goals per launch, block-sharded over %d GPU(s)"
                                   % (args.queries, B, world), "seed": args.seed},
            "gpu_launches": int(launches), "waypoints_last_share": nwp}))
    return 0


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        sys.exit(run_reference(a))
    sys.exit(run_queries(a) if a.workload == "queries2048" else run_b200(a))
