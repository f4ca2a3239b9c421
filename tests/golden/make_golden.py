#!/usr/bin/env python
"""Independent NumPy restatement of one SPH::Run step, used to mint the golden file tests/golden/c1_small.npz.

Written from the reference shaders (paths relative to /root/reference), NOT from oracle/pbf_oracle.c, so that the two
restatements check each other: a transcription error would have to be made twice, identically.  Integer results
(keys, cell starts, run lengths) must agree exactly with the C oracle; float results agree to rounding (this file
sums with NumPy's pairwise float32 sums, the oracle sums sequentially in shader order).

    python tests/golden/make_golden.py        # rewrites c1_small.npz next to this file
The initial state comes from the seeded generator (pbf_b200.dam_break == oracle.dam_break, compared bit for bit in
the test suite); everything after it is computed here.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

F = np.float32
H = F(2.0)                                      # src/SPH.cpp:58
GRID = (128, 64, 128)                           # src/SPH.h:40
WALL = np.array([16, 0, 16], np.float32)        # shaders/sph/updatepos.glsl:98
P = dict(inv_rho0=F(1.0), eps=F(5.0), gravity=F(10.0), dt=F(0.016), k=F(0.1), c_xsph=F(0.01), eps_v=F(5.0))   # src/SPH.cpp:137-144


def wpoly6(r):                                  # shaders/sph/calclambda.glsl:41-47
    tmp = H * H - r * r
    w = F(1.56668147106) * tmp * tmp * tmp / (H ** 9)
    return np.where(r > H, F(0), w).astype(np.float32)


P["scale"] = F(1.0) / wpoly6(F(0.2))            # src/SPH.cpp:142


def grad_wspiky(d):                             # calclambda.glsl:57-64, d: (m,3)
    l = np.sqrt((d * d).sum(1, dtype=np.float32))
    tmp = H - l
    with np.errstate(divide="ignore", invalid="ignore"):
        g = (F(-3 * 4.774648292756860) * tmp * tmp)[:, None] * d / (l * H ** 6)[:, None]
    g[(l > H) | (l == 0)] = 0
    return g.astype(np.float32)


def step(pos, vel, iters, ref_quirks=1, vorticity=True, extforce=False, highlight=None):
    n = pos.shape[0]
    gx, gy, gz = GRID
    # predictpos.glsl:18-38
    v = vel[:, :3].copy()
    if extforce:                                 # :27-28, applied before gravity, to particles with z > GRID_SIZE.z / 2
        far = pos[:, 2] > F(gz) / F(2)           # GRID_SIZE is injected as a vec3 (src/SPH.cpp:29): float division
        v[far, 2] += F(2) * P["gravity"] * F(-1) * P["dt"]
    v[:, 1] += P["gravity"] * F(-1) * P["dt"]
    p = (pos[:, :3] + P["dt"] * v).astype(np.float32)
    # counting.glsl:53-57
    cell = np.clip(p, 0, np.array(GRID, np.float32)).astype(np.int64)
    key = cell[:, 0] + cell[:, 2] * gx + cell[:, 1] * gx * gz
    nbits = int(gx * gy * gz - 1).bit_length()
    mask = (1 << (2 * ((nbits + 1) // 2))) - 1   # src/RadixSort.cpp:127: only these bits are sorted
    order = np.argsort(key & mask, kind="stable")
    sp, sid, skey = p[order], order.astype(np.int32), key[order]
    # findcells.glsl:34-53 (+ clear, src/NeighbourCellFinder.cpp:116-126)
    start = np.full(gx * gy * gz, -1, np.int32)
    end = np.zeros(gx * gy * gz, np.int32)
    scell = cell[order]

    def in_image(c):                             # imageStore outside the 3-D image is dropped (GL 4.3, 8.26)
        return 0 <= c[0] < gx and 0 <= c[1] < gy and 0 <= c[2] < gz

    if ref_quirks:
        start[0] = 0
    elif in_image(scell[0]):
        start[skey[0]] = 0
    for i in range(1, n):
        if tuple(scell[i]) != tuple(scell[i - 1]):
            if in_image(scell[i]):
                start[skey[i]] = i
            if in_image(scell[i - 1]):
                end[skey[i - 1]] = i
    if in_image(scell[n - 1]):
        end[skey[n - 1]] = n                     # restatement policy: the reference leaves this one stale
    # neighbourcells.glsl:52-91
    g3 = sp.astype(np.int64)                     # ivec3(pos): truncation, not clamped
    run_start = np.full((n, 9), -1, np.int32)
    run_count = np.zeros((n, 9), np.int32)
    offs = [(dy, dz) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]
    for i in range(n):
        for o, (dy, dz) in enumerate(offs):
            c, cnt = -1, 0
            for j in (-1, 0, 1):
                x, y, z = g3[i, 0] + j, g3[i, 1] + dy, g3[i, 2] + dz
                s = start[x + z * gx + y * gx * gz] if (0 <= x < gx and 0 <= y < gy and 0 <= z < gz) else -1
                if c == -1:
                    c = s
                if s != -1:
                    cnt += end[x + z * gx + y * gx * gz] - s
            run_start[i, o], run_count[i, o] = c, (cnt if c != -1 else 0)

    def neighbours(i):                           # foreachneighbour.glsl:1-10
        idx = np.concatenate([np.arange(run_start[i, o], run_start[i, o] + run_count[i, o]) for o in range(9)] + [np.zeros(0, np.int64)])
        return idx[idx != i].astype(np.int64)

    nb = [neighbours(i) for i in range(n)]
    hl = None
    if highlight is not None:                    # clearhighlight.glsl (flag &= 1), then highlight.glsl:17-30 (src/SPH.cpp:288-296)
        hl = highlight & np.uint32(1)
        for i in range(n):
            if hl[sid[i]] & 1:
                hl[sid[nb[i]]] |= np.uint32(2)
    lam = np.zeros(n, np.float32)
    for _ in range(iters):
        for i in range(n):                       # calclambda.glsl:66-103
            d = sp[i] - sp[nb[i]]
            rho = wpoly6(np.sqrt((d * d).sum(1, dtype=np.float32))).sum(dtype=np.float32)
            g = grad_wspiky(d) * P["inv_rho0"]
            s = (g * g).sum(dtype=np.float32) + (g.sum(0, dtype=np.float32) ** 2).sum(dtype=np.float32)
            lam[i] = -(rho * P["inv_rho0"] - F(1)) / (s + P["eps"])
        new = sp.copy()
        for i in range(n):                       # updatepos.glsl:43-105, Jacobi
            d = sp[i] - sp[nb[i]]
            sc = P["scale"] * wpoly6(np.sqrt((d * d).sum(1, dtype=np.float32)))
            sc = sc * sc
            sc = -P["k"] * (sc * sc)
            dp = ((lam[i] + lam[nb[i]] + sc)[:, None] * grad_wspiky(d)).sum(0, dtype=np.float32)
            new[i] = np.clip(sp[i] + P["inv_rho0"] * dp, WALL, np.array(GRID, np.float32) - WALL)
        sp = new
    # update.glsl:16-28
    pos1, vel1 = pos.copy(), vel.copy()
    vnew = ((sp - pos[sid, :3]) / P["dt"]).astype(np.float32)
    pos1[sid, :3] = sp
    vel1[sid, :3] = vnew
    if vorticity:                                # vorticity.glsl:34-86, Jacobi two-phase
        vs = vnew                                # velocity by sorted index
        om = np.zeros((n, 3), np.float32)
        vx = np.zeros((n, 3), np.float32)
        for i in range(n):
            vij = vs[nb[i]] - vs[i]
            pij = sp[i] - sp[nb[i]]
            w = wpoly6(np.sqrt((pij * pij).sum(1, dtype=np.float32)))
            vx[i] = vs[i] + P["c_xsph"] * (vij * w[:, None]).sum(0, dtype=np.float32)
            om[i] = np.cross(vij, grad_wspiky(pij)).sum(0, dtype=np.float32)
        mag = np.sqrt((om * om).sum(1, dtype=np.float32))
        for i in range(n):
            pij = sp[i] - sp[nb[i]]
            gv = (mag[nb[i]][:, None] * grad_wspiky(pij)).sum(0, dtype=np.float32)
            l = np.sqrt((gv * gv).sum(dtype=np.float32))
            if l > 0:
                gv = gv / l
            vel1[sid[i], :3] = vx[i] + P["dt"] * P["eps_v"] * np.cross(gv, om[i])
    out = dict(skey=skey.astype(np.uint32), perm=sid.astype(np.uint32), start=start, run_count=run_count, lam=lam,
               pos1=pos1, vel1=vel1)
    if hl is not None:
        out["highlight1"] = hl
    return out


def edge_scene(seed=777, n=512):
    """A small lattice block, a sparse scatter, particles outside the grid on all six sides, one that lands on y = gy
    exactly, one on x = gx exactly, and a pair of exact duplicates (all in the reference's default grid)."""
    rng = np.random.default_rng(seed)
    side = 6
    ii = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 3)
    block = 50.0 + 0.94 * ii + rng.uniform(-0.005, 0.005, (side ** 3, 3))
    m = n - side ** 3
    g = np.array(GRID, np.float64)
    rest = rng.uniform([17, 1, 17], [g[0] - 17, g[1] - 1, g[2] - 17], (m, 3))
    for k in range(48):                          # eight per face
        a = k % 3
        rest[k, a] = -rng.uniform(0.1, 3.0) if (k // 3) % 2 == 0 else g[a] + rng.uniform(0.1, 3.0)
    xyz = np.concatenate([block, rest])
    pos = np.zeros((n, 4), np.float32); pos[:, :3] = xyz
    vel = np.zeros((n, 4), np.float32); vel[:, :3] = rng.normal(0, 2.0, (n, 3))
    vel[side ** 3: side ** 3 + 48, :3] = rng.normal(0, 20.0, (48, 3))
    a = side ** 3 + 48
    pos[a] = (60.0, GRID[1], 60.0, 0.0); vel[a] = (0.0, P["gravity"] * P["dt"], 0.0, 0.0)       # p*.y = gy exactly
    pos[a + 1] = (GRID[0], 20.0, 60.0, 0.0); vel[a + 1] = (0.0, 0.0, 0.0, 0.0)                   # p*.x = gx exactly
    pos[a + 2] = pos[3]; vel[a + 2] = vel[3]                                                       # exact duplicate
    return pos, vel


def main():
    import pbf_b200
    n3, seed, iters = (8, 8, 8), 4242, 3
    if "--edge" not in sys.argv:
        pos, vel = pbf_b200.dam_break(*n3, seed=seed)
        out = step(pos, vel, iters)
        out.pop("perm")                          # c1_small.npz predates the permutation entry; keep the file as committed
        np.savez_compressed(os.path.join(HERE, "c1_small.npz"), n3=np.array(n3), grid=np.array(GRID), seed=seed, iters=iters,
                            ref_quirks=1, pos0=pos, **out)
        print("wrote c1_small.npz:", {k: v.shape for k, v in out.items()})
    # edge cases: out-of-grid particles, the ceiling and x = gx planes, duplicates -- in both quirk modes
    pos, vel = edge_scene()
    for quirks in (1, 0):
        out = step(pos, vel, 2, ref_quirks=quirks)
        name = "edge_small_q%d.npz" % quirks
        np.savez_compressed(os.path.join(HERE, name), grid=np.array(GRID), iters=2, ref_quirks=quirks, pos0=pos, vel0=vel, **out)
        print("wrote %s:" % name, {k: v.shape for k, v in out.items()})
    # external force (F key, src/Simulation.cpp:280-282) and highlight marks (H + click, :160-195) on a lattice block
    pos, vel = pbf_b200.dam_break(8, 8, 8, origin=(60.5, 0.5, 60.5), seed=99)      # straddles z = 64
    hl = np.zeros(pos.shape[0], np.uint32)
    hl[[3, 100, 400]] = 1
    hl[[7, 8]] = 2                               # stale neighbour marks from the step before
    hl[200] = 3
    out = step(pos, vel, 2, extforce=True, highlight=hl)
    np.savez_compressed(os.path.join(HERE, "force_highlight_small.npz"), grid=np.array(GRID), iters=2, ref_quirks=1, pos0=pos,
                        vel0=vel, highlight0=hl, **out)
    print("wrote force_highlight_small.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
