"""Domain decomposition on the GPU: k logical strips on one device must reproduce the
single-grid solve ('multi-GPU without a cluster')."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [2, 3, 4])
def test_k_strips_equal_single_grid(pkg, k):
    import torch
    sh, api, syn = pkg.sharding, pkg.cuda_api, pkg.synthetic
    ny, nx = 384, 320
    cost = syn.smooth_cost_map(ny, nx, seed=21, obstacle_fraction=0.04)
    gi, gj = syn.free_interior_cell_near(cost <= 0, 200, 40)
    whole = api.DeviceLayer(nx, ny)
    whole.set_cost_map(cost)
    whole.solve_total_cost([(gi, gj)])
    T1 = whole.download_total_cost()
    strips = []
    for r in range(k):
        lay = sh.StripLayout(ny, k, r)
        strips.append(sh.CudaStrip(api, lay, nx, cost[lay.r0:lay.r1], 0, torch))
    rounds = sh.dd_solve_lockstep(strips, (gi, gj))
    Tk = np.vstack([s.own_rows() for s in strips])
    assert rounds >= 2
    assert np.array_equal(np.isinf(Tk), np.isinf(T1))
    assert rel_err(Tk, T1) <= 1e-12


def test_batched_slots_resume_and_rows(pkg):
    """export/import round trip and resume after lowering a row by hand."""
    import torch
    api = pkg.cuda_api
    n = 128
    dev = api.DeviceLayer(n, n)
    dev.set_cost_map(np.ones((n, n)))
    dev.solve_total_cost([(64, 64)])
    T0 = dev.download_total_cost()
    row = torch.empty(n, dtype=torch.float64, device="cuda")
    dev.export_rows(10, 1, row.data_ptr(), True)
    assert np.array_equal(row.cpu().numpy(), T0[10])
    lower = torch.full((n,), 0.5, dtype=torch.float64, device="cuda")
    assert dev.import_rows_min(10, 1, lower.data_ptr(), True)
    assert not dev.import_rows_min(10, 1, lower.data_ptr(), True)
    dev.solve_resume([(9, 12)])
    T1 = dev.download_total_cost()
    assert np.all(T1 <= T0) and T1[11, 5] == 1.5 and T1[10, 5] == 0.5


@pytest.mark.parametrize("phases", [1, 5, 40])
def test_phase_bounded_solve_equals_one_shot(pkg, phases):
    """dymu_solve_start + dymu_solve_advance in slices of `phases` solver phases reach the same
    fixed point as one dymu_solve_total_cost launch (the pending work lists survive between
    launches)."""
    api, syn = pkg.cuda_api, pkg.synthetic
    ny, nx = 416, 352
    cost = syn.smooth_cost_map(ny, nx, seed=8, obstacle_fraction=0.05)
    goal = syn.free_interior_cell_near(cost <= 0, 100, 300)
    dev = api.DeviceLayer(nx, ny)
    dev.set_cost_map(cost)
    one = dev.solve_total_cost([goal])
    T1 = dev.download_total_cost()
    st = dev.solve_start(goal, phases)
    launches, total_phases = 1, st["outer_iterations"]
    while not st["converged"]:
        assert st["outer_iterations"] == phases
        st = dev.solve_advance([], 0.0, phases)
        launches += 1
        total_phases += st["outer_iterations"]
        assert launches < 100000
    Tb = dev.download_total_cost()
    assert np.array_equal(np.isinf(Tb), np.isinf(T1))
    assert rel_err(Tb, T1) <= 1e-12
    if phases == 1:
        assert launches > 10
    # slicing must not change the schedule much: same order of magnitude of phases
    assert total_phases <= 2 * one["outer_iterations"] + 4
    # nothing pending any more: another advance is a no-op
    again = dev.solve_advance([], 0.0, phases)
    assert again["converged"] and again["tile_activations"] == 0


@pytest.mark.parametrize("k,phases", [(2, 3), (3, 8), (4, 2)])
def test_pipelined_strips_equal_single_grid(pkg, k, phases):
    import torch
    sh, api, syn = pkg.sharding, pkg.cuda_api, pkg.synthetic
    ny, nx = 384, 320
    cost = syn.smooth_cost_map(ny, nx, seed=21, obstacle_fraction=0.04)
    gi, gj = syn.free_interior_cell_near(cost <= 0, 200, 180)   # goal in an interior strip for k >= 3
    whole = api.DeviceLayer(nx, ny)
    whole.set_cost_map(cost)
    whole.solve_total_cost([(gi, gj)])
    T1 = whole.download_total_cost()
    strips = []
    for r in range(k):
        lay = sh.StripLayout(ny, k, r)
        strips.append(sh.CudaStrip(api, lay, nx, cost[lay.r0:lay.r1], 0, torch))
    rounds = sh.dd_solve_lockstep_pipelined(strips, (gi, gj), phases_per_round=phases)
    Tk = np.vstack([s.own_rows() for s in strips])
    assert rounds >= 2
    assert np.array_equal(np.isinf(Tk), np.isinf(T1))
    assert rel_err(Tk, T1) <= 1e-12
