"""Model check of the rule by which the solve kernel delivers a tile of the total-cost matrix
before the solve has finished (dymu_set_total_cost_export; k_fim in
planning-path_planning_b200/csrc/dymu_fim.cu: Params::tmax).

The kernel keeps, per tile, an upper bound of the tile's values as of its last write-back (+inf
while a passable cell is unreached) and, per pending tile, a key = the smallest changed value it
was woken for.  At the end of a phase every tile whose bound lies below the smallest key that was
pending when the phase began is stored to the caller's matrix.  The claim: such a tile never
changes again (the update is upwind -- a change can only produce values larger than itself).

The model is a small tile fast-iterative solver with the kernel's scheduling rules (priority band,
deferral, wake-up keys, wake-up filter against the facing halo cell, tiles of one phase seeing
each other's values in arbitrary order) on random maps with obstacles and enclosed pockets; every
tile snapshot taken by the delivery rule must equal the converged map.  CPU only."""
import math
import random

import numpy as np

INF = float("inf")
TILE = 4


def eikonal(tx, ty, c):
    """propagateGlobalNode, src/DyMu_GlobalPathPlanning.cpp:527-535"""
    tx, ty, c = float(tx), float(ty), float(c)  # (plain floats: inf - inf is a quiet NaN)
    d = tx - ty
    if abs(d) < c and tx < INF and ty < INF:
        return (tx + ty + math.sqrt(2 * c * c - d * d)) / 2
    return min(tx, ty) + c


def relax_tile(T, C, y0, x0):
    """Gauss-Seidel sweeps of one tile against the current plane until nothing changes.
    Returns the set of changed cells."""
    ny, nx = T.shape
    changed = set()
    again = True
    while again:
        again = False
        for y in range(y0, y0 + TILE):
            for x in range(x0, x0 + TILE):
                c = C[y, x]
                if not c < INF:
                    continue
                l = T[y, x - 1] if x > 0 else INF
                r = T[y, x + 1] if x + 1 < nx else INF
                u = T[y - 1, x] if y > 0 else INF
                d = T[y + 1, x] if y + 1 < ny else INF
                t = eikonal(min(l, r), min(u, d), c)
                if t < T[y, x]:
                    T[y, x] = t
                    changed.add((y, x))
                    again = True
    return changed


def solve_with_delivery(rng, C, goal, band):
    ny, nx = C.shape
    nty, ntx = ny // TILE, nx // TILE
    T = np.full((ny, nx), INF)
    T[goal] = 0.0
    key = {(goal[0] // TILE, goal[1] // TILE): 0.0}   # pending tiles and their keys
    tmax = {}                                        # bound per tile as of its last write-back
    delivered = {}                                   # tile -> snapshot
    phases = 0
    while key:
        phases += 1
        assert phases < 10000
        gm = min(key.values())
        limit = gm + band
        todo = [t for t, k in key.items() if k <= limit]
        rng.shuffle(todo)  # tiles of one phase see each other's values in arbitrary order
        nxt = {t: k for t, k in key.items() if k > limit}  # deferred with their keys
        for (ty, tx) in todo:
            y0, x0 = ty * TILE, tx * TILE
            halo_before = T.copy()  # what this activation staged (the kernel loads the halo once)
            changed = relax_tile(T, C, y0, x0)
            # bound of the tile after the write-back
            vals = [T[y, x] for y in range(y0, y0 + TILE) for x in range(x0, x0 + TILE) if C[y, x] < INF]
            tmax[(ty, tx)] = max(vals) if vals else 0.0
            # wake the neighbours whose halo went stale: only across an edge where a changed cell
            # dropped below the cell facing it (as staged), with the smallest such value as key
            for (dy, dx) in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                nty_, ntx_ = ty + dy, tx + dx
                if not (0 <= nty_ < nty and 0 <= ntx_ < ntx):
                    continue
                k = INF
                for (y, x) in changed:
                    fy, fx = y + dy, x + dx
                    if (fy // TILE, fx // TILE) == (nty_, ntx_) and T[y, x] < halo_before[fy, fx]:
                        k = min(k, T[y, x])
                if k < INF:
                    nxt[(nty_, ntx_)] = min(nxt.get((nty_, ntx_), INF), k)
        # ---- the delivery rule, with the minimum key of the START of the phase
        for t, b in tmax.items():
            if b < gm and t not in delivered:
                y0, x0 = t[0] * TILE, t[1] * TILE
                delivered[t] = T[y0:y0 + TILE, x0:x0 + TILE].copy()
        key = nxt
    return T, delivered, phases


def random_map(rng, ny, nx):
    C = np.array([[rng.choice((1.0, 1.0, 1.5, 3.0, 7.0)) for _ in range(nx)] for _ in range(ny)])
    for _ in range(rng.randrange(0, 6)):  # obstacle bars, some of them closing pockets
        y, x = rng.randrange(ny), rng.randrange(nx)
        if rng.random() < 0.5:
            C[y, x:min(nx, x + rng.randrange(2, 9))] = INF
        else:
            C[y:min(ny, y + rng.randrange(2, 9)), x] = INF
    if rng.random() < 0.5:  # a closed box: passable cells the wave never reaches
        y, x = rng.randrange(1, ny - 4), rng.randrange(1, nx - 4)
        C[y - 1:y + 3, x - 1:x + 3] = INF
        C[y:y + 2, x:x + 2] = 1.0
    return C


def test_every_tile_delivered_early_is_final():
    rng = random.Random(4096)
    n_early = 0
    for trial in range(60):
        ny, nx = rng.choice((12, 16, 20)), rng.choice((12, 16, 24))
        C = random_map(rng, ny, nx)
        free = [(y, x) for y in range(ny) for x in range(nx) if C[y, x] < INF]
        goal = rng.choice(free)
        band = rng.choice((0.0, 2.0, 8.0, INF))
        T, delivered, phases = solve_with_delivery(rng, C, goal, band)
        # the converged plane is a fixed point of the update
        for (y, x) in free:
            if (y, x) == goal:
                continue
            l = T[y, x - 1] if x > 0 else INF
            r = T[y, x + 1] if x + 1 < nx else INF
            u = T[y - 1, x] if y > 0 else INF
            d = T[y + 1, x] if y + 1 < ny else INF
            assert not eikonal(min(l, r), min(u, d), C[y, x]) < T[y, x]
        for (ty, tx), snap in delivered.items():
            final = T[ty * TILE:(ty + 1) * TILE, tx * TILE:(tx + 1) * TILE]
            assert np.array_equal(snap, final), (trial, ty, tx, band)
        n_early += len(delivered)
    assert n_early > 200  # the rule does fire
