"""Seeded synthetic inputs of the shapes BASELINE.json names.

Mars-like DEM: spectral-synthesis fBm elevation (power spectrum ~ k^-beta)
scaled to a target RMS slope, plus craters with raised rims; terrain classes
1..4 from an independent, smoother fBm; a cost look-up table
[terrain][locomotion][slope] = base_t * (1 + slope_deg / 15).  Obstacles arise
inside DyMu's own computeCostMap where the slope exceeds the table's maximum.
All randomness comes from numpy.random.default_rng(seed).
"""
import numpy as np

DEFAULT_SEED = 20261018


def fbm(ny, nx, beta, rng):
    """Real 2-D field with isotropic power spectrum ~ k^-beta, zero mean, unit std."""
    ky = np.fft.fftfreq(ny)[:, None]
    kx = np.fft.rfftfreq(nx)[None, :]
    k = np.sqrt(kx * kx + ky * ky)
    k[0, 0] = 1.0
    amp = k ** (-beta / 2.0)
    amp[0, 0] = 0.0
    phase = rng.uniform(0.0, 2.0 * np.pi, size=amp.shape)
    spec = amp * np.exp(1j * phase)
    f = np.fft.irfft2(spec, s=(ny, nx))
    f -= f.mean()
    f /= f.std()
    return f


def _add_craters(elev, n_craters, rng, rmin=4.0, rmax=40.0):
    ny, nx = elev.shape
    for _ in range(n_craters):
        r = rng.uniform(rmin, min(rmax, min(nx, ny) / 6.0))
        cx = rng.uniform(0, nx)
        cy = rng.uniform(0, ny)
        R = int(np.ceil(1.6 * r))
        x0, x1 = max(0, int(cx) - R), min(nx, int(cx) + R + 1)
        y0, y1 = max(0, int(cy) - R), min(ny, int(cy) + R + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        yy, xx = np.mgrid[y0:y1, x0:x1]
        d = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2) / r
        depth = rng.uniform(0.13, 0.2) * r
        bowl = np.where(d < 1.0, -depth * (1.0 - d * d), 0.0)
        rim = 0.25 * depth * np.exp(-((d - 1.0) / 0.18) ** 2)
        elev[y0:y1, x0:x1] += bowl + rim


def default_lut(n_loc=1):
    """cost_data, slope_values, locomotion mode names (computeCostMap inputs, G.cpp:128-149)."""
    slopes = np.arange(0.0, 30.0 + 1e-9, 5.0)
    base = np.array([0.0, 1.0, 1.5, 2.5, 4.0])
    rows = []
    for t in range(5):
        for l in range(n_loc):
            b = base[t] if t > 0 else 10.0  # terrain 0 row is never looked up (obstacle)
            rows.append(b * (1.0 + 0.15 * l) * (1.0 + slopes / 15.0))
    locs = ["DRIVING", "WHEEL_WALKING", "CRABBING"][:n_loc]
    return np.concatenate(rows), slopes, locs


def mars_dem(ny, nx, seed=DEFAULT_SEED, rms_slope_deg=8.0, craters=None):
    """Returns (elevation[ny,nx] float64, terrain[ny,nx] float64 with classes 1..4)."""
    rng = np.random.default_rng(seed)
    elev = fbm(ny, nx, 2.2, rng)
    gy, gx = np.gradient(elev)
    rms = np.sqrt(np.mean(gx * gx + gy * gy))
    elev *= np.tan(np.deg2rad(rms_slope_deg)) / rms
    if craters is None:
        # constant areal density (a count that grows only with the edge length would let the
        # obstacle fraction fall with the map size)
        craters = max(1, int(round(0.85 * nx * ny / 4096.0)))
    _add_craters(elev, craters, rng)
    tfield = fbm(ny, nx, 3.0, rng)
    q = np.quantile(tfield, [0.4, 0.7, 0.9])
    terrain = 1.0 + np.digitize(tfield, q).astype(np.float64)
    return np.ascontiguousarray(elev), np.ascontiguousarray(terrain)


def smooth_cost_map(ny, nx, seed=DEFAULT_SEED, obstacle_fraction=0.03):
    """A direct setCostMap input: smooth cost in [0.05, 2.6] with constant-coverage discs
    of obstacles (cost 0), as in the survey probe."""
    rng = np.random.default_rng(seed)
    f = fbm(ny, nx, 2.6, rng)
    f = (f - f.min()) / (f.max() - f.min())
    cost = 0.05 + 2.55 * f
    area = obstacle_fraction * nx * ny
    yy, xx = np.mgrid[0:ny, 0:nx]
    placed = 0.0
    while placed < area:
        r = rng.uniform(2.0, max(3.0, min(nx, ny) / 40.0))
        cx, cy = rng.uniform(0, nx), rng.uniform(0, ny)
        x0, x1 = max(0, int(cx - r) - 1), min(nx, int(cx + r) + 2)
        y0, y1 = max(0, int(cy - r) - 1), min(ny, int(cy + r) + 2)
        m = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r
        cost[y0:y1, x0:x1][m] = 0.0
        placed += np.pi * r * r
    return np.ascontiguousarray(cost)


def free_interior_cell_near(obstacle, ci, cj, margin=2):
    """Nearest (i, j) to (ci, cj) whose 3x3 neighbourhood is obstacle-free and that lies
    `margin` cells inside the map (valid goal for setGoal, G.cpp:322-357)."""
    ny, nx = obstacle.shape
    ob = obstacle.astype(bool)
    blocked = ob.copy()
    blocked[1:, :] |= ob[:-1, :]
    blocked[:-1, :] |= ob[1:, :]
    blocked[:, 1:] |= ob[:, :-1]
    blocked[:, :-1] |= ob[:, 1:]
    blocked[1:, 1:] |= ob[:-1, :-1]
    blocked[:-1, :-1] |= ob[1:, 1:]
    blocked[1:, :-1] |= ob[:-1, 1:]
    blocked[:-1, 1:] |= ob[1:, :-1]
    blocked[:margin, :] = True
    blocked[-margin:, :] = True
    blocked[:, :margin] = True
    blocked[:, -margin:] = True
    jj, ii = np.nonzero(~blocked)
    k = np.argmin((ii - ci) ** 2 + (jj - cj) ** 2)
    return int(ii[k]), int(jj[k])


def obstacle_frame(h, w, res, centre_xy, discs):
    """uint8 traversability image (image convention, Y down; L.cpp:225-238) centred on
    centre_xy with obstacle discs [(x, y, radius_m), ...] in world coordinates."""
    img = np.zeros((h, w), dtype=np.uint8)
    ox = centre_xy[0] - res * w / 2.0
    oy = centre_xy[1] + res * h / 2.0
    jj, ii = np.mgrid[0:h, 0:w]
    px = ox + ii * res
    py = oy - jj * res
    for (x, y, r) in discs:
        img[(px - x) ** 2 + (py - y) ** 2 <= r * r] = 1
    return img
