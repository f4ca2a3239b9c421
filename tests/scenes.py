"""Seeded edge-case scenes shared by the CPU (oracle-only) and GPU (parity) tests.  Every scene is a multiple of 512
particles (src/Simulation.cpp:202)."""
import numpy as np


def _pad4(xyz):
    out = np.zeros((xyz.shape[0], 4), np.float32)
    out[:, :3] = xyz
    return out


def sparse_gas(n=4096, grid=(128, 64, 128), seed=11, speed=3.0):
    """Uniformly random particles inside the walls: ~0.006 particles per cell, so 256 consecutive sorted particles span
    hundreds of cell rows, almost every neighbour run is empty and most cells are."""
    rng = np.random.default_rng(seed)
    lo = np.array([16.0, 0.0, 16.0]); hi = np.array([grid[0] - 16.0, grid[1], grid[2] - 16.0])
    pos = _pad4(rng.uniform(lo, hi, (n, 3)))
    vel = _pad4(rng.normal(0, speed, (n, 3)))
    return pos, vel


def clump(n=2048, cells=4, origin=(40.0, 10.0, 40.0), seed=12):
    """n particles inside cells^3 unit cells (32 per cell at the defaults): merged 3-cell runs hold ~96 candidates, above
    the 31 the packed per-tile plan can describe, so these tiles must fall back to the general path."""
    rng = np.random.default_rng(seed)
    pos = _pad4(np.asarray(origin) + rng.uniform(0.0, cells, (n, 3)))
    vel = _pad4(np.zeros((n, 3)))
    return pos, vel


def escapees(n=1024, grid=(128, 64, 128), seed=13):
    """A lattice block plus particles outside the grid on every side and one that lands on y = gy exactly (key bit above
    the sorted bits, no cell: SURVEY.md 8c-v); velocities throw more particles out of the grid during predict."""
    rng = np.random.default_rng(seed)
    side = 8
    m = side ** 3
    ii = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 3)
    block = 50.0 + 0.94 * ii + rng.uniform(-0.005, 0.005, (m, 3))
    extra = n - m
    out = rng.uniform(-3.0, 3.0, (extra, 3))
    which = rng.integers(0, 6, extra)
    g = np.asarray(grid, np.float64)
    base = rng.uniform([16, 1, 16], [g[0] - 16, g[1] - 1, g[2] - 16], (extra, 3))
    for k in range(extra):
        a = which[k] % 3
        base[k, a] = -abs(out[k, a]) - 0.01 if which[k] < 3 else g[a] + abs(out[k, a]) + 0.01
    pos = _pad4(np.concatenate([block, base]))
    vel = _pad4(np.concatenate([rng.normal(0, 1.0, (m, 3)), rng.normal(0, 40.0, (extra, 3))]))
    # ceiling case: v.y + (-g dt) = 0 exactly, so p*.y = 64.0 = gy
    pos[m] = (60.0, float(grid[1]), 60.0, 0.0)
    vel[m] = (0.0, np.float32(10.0) * np.float32(0.016), 0.0, 0.0)
    return pos, vel


def splash(n3=(16, 16, 32), grid=(128, 64, 128), seed=14):
    """Two lattice blocks thrown at each other (scaled down): heavy motion, wall hits."""
    import oracle
    a, _ = oracle.dam_break(*n3, origin=(20.5, 4.5, 20.5), seed=seed)
    b, _ = oracle.dam_break(*n3, origin=(70.5, 6.5, 60.5), seed=seed + 1, id0=a.shape[0])
    pos = np.concatenate([a, b])
    vel = np.zeros_like(pos)
    vel[: a.shape[0], :3] = (25.0, 10.0, 20.0)
    vel[a.shape[0]:, :3] = (-25.0, 5.0, -20.0)
    return pos, vel
