"""Domain-decomposed total-cost solve of one large grid across the GPUs of a node
(BASELINE.json configs[4]).  Launch:  torchrun --nnodes=1 --nproc-per-node N
--master-addr 127.0.0.1 scripts/dd_solve.py --size 16384 [--verify]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dymu_b200

ap = argparse.ArgumentParser()
ap.add_argument("--size", dest="n", type=int, default=16384)
ap.add_argument("--base", type=int, default=4096, help="edge of the periodic fBm cost tile")
ap.add_argument("--verify", action="store_true")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--phases", type=int, default=0,
                help="> 0: pipelined exchange every that many solver phases (dd_solve_pipelined)")
a = ap.parse_args()

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = dymu_b200.load()
sh, api, syn = pkg.sharding, pkg.cuda_api, pkg.synthetic
n, base = a.n, min(a.base, a.n)
tile = syn.smooth_cost_map(base, base, seed=20261018, obstacle_fraction=0.03)   # periodic (FFT fBm)
reps_y = n // base
lay = sh.StripLayout(n, world, rank)
rows = np.arange(lay.r0, lay.r1) % base
cost_own = np.tile(tile[rows], (1, reps_y))
goal = syn.free_interior_cell_near(tile <= 0, base // 2, base // 2)
goal = (goal[0] + (reps_y // 2) * base, goal[1] + (reps_y // 2) * base)
strip = sh.CudaStrip(api, lay, n, cost_own, local, torch)
comm = sh.TorchComm(rank, world, torch.device("cuda", local)) if world > 1 else None
times, rounds = [], 0
for r in range(a.reps):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    if world > 1:
        rounds = sh.dd_solve_pipelined(strip, comm, goal, a.phases) if a.phases > 0 \
            else sh.dd_solve(strip, comm, goal)
    else:
        strip.start(goal)
        rounds = 1
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    times.append(time.perf_counter() - t0)
kernel_ms = sum(s["kernel_ms"] for s in strip.stats[-max(1, len(strip.stats) // a.reps):])
t = torch.tensor([min(times), kernel_ms], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ok = None
if a.verify:
    whole = api.DeviceLayer(n, n, device=local)
    whole.set_cost_map(np.tile(tile, (reps_y, reps_y)))
    whole.solve_total_cost([goal])
    T1 = whole.download_total_cost()[lay.r0:lay.r1]
    Tk = strip.own_rows()
    fin = np.isfinite(T1) & (T1 > 0)
    err = float(np.max(np.abs(Tk[fin] - T1[fin]) / T1[fin])) if fin.any() else 0.0
    good = torch.tensor([1 if (np.array_equal(np.isinf(Tk), np.isinf(T1)) and err <= 1e-12) else 0],
                        device="cuda")
    if world > 1:
        dist.all_reduce(good, op=dist.ReduceOp.MIN)
    ok = bool(good.item())
if rank == 0 and getattr(strip, "laps", None):
    print("host wall per stage, last solve (ms):", {k: round(v * 1e3, 2) for k, v in strip.laps.items()},
          file=sys.stderr)
if rank == 0:
    print(json.dumps({"workload": "%dx%d single grid, row strips" % (n, n), "n_gpus": world,
                      "wall_ms": float(t[0]) * 1e3, "max_rank_kernel_ms": float(t[1]),
                      "exchange_rounds": rounds, "phases_per_round": a.phases, "verified": ok}))
if world > 1:
    dist.destroy_process_group()
