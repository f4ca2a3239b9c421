"""Edge cases of the CUDA path against the oracle: sparse and over-dense scenes (general sweep path), particles outside
the grid / on the ceiling plane (keys without a cell), fast splashes, the smallest handle, determinism, and the two
sweep paths against each other.  Integer results bit-exact, floats within the north star's tolerance (scaled by the
displacement where a scene is violent by construction)."""
import os

import numpy as np
import pytest

import oracle
import pbf_b200
import scenes

pytestmark = pytest.mark.gpu

GRID = (128, 64, 128)
POS_TOL = 1e-5 * 128.0


def oracle_params(sph):
    P = oracle.default_params()
    p = sph._get()
    for k in ("one_over_rho_0", "epsilon", "gravity", "timestep", "tensile_instability_k",
              "tensile_instability_scale", "xsph_viscosity_c", "vorticity_epsilon"):
        setattr(P, k, getattr(p, k))
    return P


def check_tables(sph, g, pos, vel, quirks):
    """predict + sort + cells on both sides: everything integer must be identical."""
    P = oracle_params(sph)
    sph.predict(); sph.sort(); sph.build_cells()
    rec, keys = sph.get_predicted()
    orec = oracle.predict(pos, vel, P, g)
    assert np.array_equal(rec.view(np.uint32), orec.view(np.uint32))
    assert np.array_equal(keys, oracle.keys(orec, g))
    skeys, perm, srec = sph.get_sorted()
    osorted, okeys = oracle.sort(orec, g)
    assert np.array_equal(skeys, okeys)
    assert np.array_equal(perm, osorted[:, 3].view(np.int32).astype(np.uint32))
    start, end = sph.get_cell_ranges()
    ostart, oend = oracle.findcells(osorted, g)
    assert np.array_equal(start, ostart)
    occ = ostart != -1
    if quirks:
        occ[0] = False
    assert np.array_equal(end[occ], oend[occ])
    rs, rc = sph.get_neighbour_runs()
    ors, orc = oracle.neighbourcells(osorted, g, ostart, oend)
    assert np.array_equal(rc, orc)
    assert np.array_equal(rs[rc > 0], ors[orc > 0])
    return osorted, ors, orc


def check_solver_stages(sph, g, cur, rs, rc, iters=2):
    """lambda and delta-p per iteration on identical inputs; tolerances relative to the size of the result."""
    P = oracle_params(sph)
    for it in range(iters):
        sph.calc_lambda()
        lam = sph.get_lambda()
        olam, _ = oracle.calclambda(cur, rs, rc, P)
        assert np.max(np.abs(lam - olam)) <= 1e-5 * max(1.0, np.max(np.abs(olam))), it
        sph.update_positions()
        _, _, new = sph.get_sorted()
        onew = oracle.updatepos(cur, rs, rc, olam, P, g)
        moved = np.max(np.abs(onew[:, :3] - cur[:, :3]))
        assert np.max(np.abs(new[:, :3] - onew[:, :3])) <= 1e-5 * max(10.0, moved), it
        cur = new


@pytest.mark.parametrize("quirks", [True, False])
@pytest.mark.parametrize("scene", ["sparse_gas", "clump", "escapees", "splash"])
def test_edge_scene_tables_and_solver(built_lib, scene, quirks):
    pos, vel = getattr(scenes, scene)()
    g = oracle.make_grid(*GRID, ref_quirks=int(quirks))
    sph = pbf_b200.SPH(pos.shape[0], GRID, ref_quirks=quirks)
    sph.upload(pos, vel)
    osorted, ors, orc = check_tables(sph, g, pos, vel, quirks)
    tiles, tiled = sph.tile_stats()
    if scene == "clump":
        assert tiled < tiles          # ~96 candidates per merged run: more than the packed plan describes (31)
    check_solver_stages(sph, g, osorted, ors, orc)


@pytest.mark.parametrize("scene", ["sparse_gas", "escapees", "splash"])
def test_edge_scene_whole_steps(built_lib, scene):
    pos, vel = getattr(scenes, scene)()
    g = oracle.make_grid(*GRID)
    sph = pbf_b200.SPH(pos.shape[0], GRID)
    sph.SetNumSolverIterations(3)
    sph.SetVorticityConfinementEnabled(True)
    sph.upload(pos, vel)
    P = oracle_params(sph)
    sim = oracle.Sim(pos.shape[0], g)
    opos, ovel = pos.copy(), vel.copy()
    for step in range(4):
        before = opos.copy()
        sph.Run()
        sim.step(opos, ovel, P, 3, vorticity=True)
        gpos, gvel = sph.download()
        scale = max(1.0, np.max(np.abs(opos - before)) / 0.5)     # escapees are hauled back by tens of cells in one step
        assert np.max(np.abs(gpos - opos)) < POS_TOL * scale, step
        assert np.max(np.abs(gvel - ovel)) < POS_TOL * scale / 0.016, step
        sph.upload(opos, ovel)


def test_smallest_handle(built_lib):
    """512 particles (one sort block, src/SPH.cpp:25) in an 8 x 8 x 8 grid with thin walls."""
    grid, wall = (8, 8, 8), (1.0, 0.0, 1.0)
    pos, vel = oracle.dam_break(8, 8, 8, origin=(1.2, 0.3, 1.2), spacing=0.7)
    g = oracle.make_grid(*grid, wall=wall)
    sph = pbf_b200.SPH(512, grid, wall=wall)
    sph.SetNumSolverIterations(2)
    sph.upload(pos, vel)
    P = oracle_params(sph)
    sim = oracle.Sim(512, g)
    opos, ovel = pos.copy(), vel.copy()
    for step in range(3):
        sph.Run()
        sim.step(opos, ovel, P, 2)
        gpos, gvel = sph.download()
        assert np.max(np.abs(gpos - opos)) < POS_TOL, step
        sph.upload(opos, ovel)
    assert gpos[:, 0].min() >= 1.0 and gpos[:, 0].max() <= 7.0 and gpos[:, 1].min() >= 0.0


def test_zero_iterations_and_zero_steps(built_lib):
    """K = 0: predict, sort, update only (the reference's loop body simply does not run, src/SPH.cpp:302-311)."""
    pos, vel = oracle.dam_break(16, 16, 16)
    g = oracle.make_grid(*GRID)
    sph = pbf_b200.SPH(pos.shape[0], GRID)
    sph.SetNumSolverIterations(0)
    sph.upload(pos, vel)
    sph.Run(0)
    p0, _ = sph.download()
    assert np.array_equal(p0, pos)
    sph.Run()
    gpos, gvel = sph.download()
    P = oracle_params(sph)
    opos, ovel = pos.copy(), vel.copy()
    oracle.Sim(pos.shape[0], g).step(opos, ovel, P, 0)
    assert np.array_equal(gpos.view(np.uint32), opos.view(np.uint32))       # no sweep ran: pure IEEE arithmetic, bit exact
    assert np.array_equal(gvel.view(np.uint32), ovel.view(np.uint32))


@pytest.mark.parametrize("graph", [False, True])
def test_deterministic(built_lib, graph):
    """Same state in, same bits out -- across handles and across the graph / direct launch paths."""
    pos, vel = scenes.splash()
    outs = []
    for use_graph in (graph, graph, not graph):
        sph = pbf_b200.SPH(pos.shape[0], GRID, use_graph=use_graph)
        sph.SetNumSolverIterations(3)
        sph.SetVorticityConfinementEnabled(True)
        sph.upload(pos, vel)
        sph.Run(5)
        outs.append(sph.download())
        sph.close()
    for p, v in outs[1:]:
        assert np.array_equal(p.view(np.uint32), outs[0][0].view(np.uint32))
        assert np.array_equal(v.view(np.uint32), outs[0][1].view(np.uint32))


def test_tiled_and_general_sweeps_agree(built_lib):
    """PBF_GENERAL_SWEEPS=1 sends every tile down the global-memory walk; both paths visit the same candidates."""
    pos, vel = oracle.dam_break(32, 32, 32)
    res = []
    for general in ("0", "1"):
        os.environ["PBF_GENERAL_SWEEPS"] = general
        try:
            sph = pbf_b200.SPH(pos.shape[0], GRID)
        finally:
            os.environ.pop("PBF_GENERAL_SWEEPS")
        sph.SetNumSolverIterations(3)
        sph.SetVorticityConfinementEnabled(True)
        sph.upload(pos, vel)
        sph.Run()
        tiles, tiled = sph.tile_stats()
        assert (tiled == 0) if general == "1" else (tiled > 0.9 * tiles)
        res.append(sph.download())
    assert np.max(np.abs(res[0][0] - res[1][0])) < 2e-5
    assert np.max(np.abs(res[0][1] - res[1][1])) < 2e-5 / 0.016


def test_step_host_matches_step(built_lib):
    """The end-to-end call (host buffers in and out, position read-back overlapped with the vorticity kernels) returns
    exactly what upload + Run + download returns."""
    import torch
    pos, vel = scenes.splash()
    for vort in (False, True):
        a = pbf_b200.SPH(pos.shape[0], GRID)
        a.SetNumSolverIterations(3)
        a.SetVorticityConfinementEnabled(vort)
        a.upload(pos, vel)
        a.Run(3)
        apos, avel = a.download()
        b = pbf_b200.SPH(pos.shape[0], GRID)
        b.SetNumSolverIterations(3)
        b.SetVorticityConfinementEnabled(vort)
        hp = torch.from_numpy(pos.copy()).pin_memory()
        hv = torch.from_numpy(vel.copy()).pin_memory()
        b.step_host(hp, hv, 2)
        b.step_host(hp, hv, 1)
        assert np.array_equal(hp.numpy().view(np.uint32), apos.view(np.uint32))
        assert np.array_equal(hv.numpy().view(np.uint32), avel.view(np.uint32))


def test_pick_and_toggle_highlight(built_lib):
    """Ray-cast picking against a NumPy ray/sphere search; the highlight toggle of Simulation::OnMouseDown."""
    pos, vel = oracle.dam_break(16, 16, 16)
    sph = pbf_b200.SPH(pos.shape[0], GRID)
    sph.upload(pos, vel)
    rng = np.random.default_rng(21)

    def reference(o, d, radius):
        d = d / np.linalg.norm(d)
        c = pos[:, :3].astype(np.float64) - o
        b = c @ d
        disc = b * b - np.einsum("ij,ij->i", c, c) + radius * radius
        t = np.where(disc >= 0, b - np.sqrt(np.maximum(disc, 0)), np.inf)
        t[t < 0] = np.inf
        return int(np.argmin(t)) if np.isfinite(t.min()) else -1, t

    hits = 0
    for k in range(40):
        o = np.array([40.0, 8.0, 10.0]) + rng.uniform(-6, 6, 3)
        target = pos[rng.integers(0, pos.shape[0]), :3] + rng.uniform(-0.3, 0.3, 3)
        d = target - o
        want, t = reference(o, d, 0.5)
        got = sph.pick_particle(o, d, 0.5)
        if got != want:             # float32 vs float64 may swap two hits at nearly the same depth
            assert got >= 0 and abs(t[got] - t[want]) < 1e-3
        hits += got >= 0
    assert hits == 40
    assert sph.pick_particle((40.0, 200.0, 40.0), (0.0, 1.0, 0.0)) == -1          # looking away from the fluid
    with pytest.raises(RuntimeError):
        sph.pick_particle((0, 0, 0), (0, 0, 0))
    sph.toggle_highlight(77)
    _, _, hl = sph.download(highlight=True)
    assert hl[77] == 1 and hl.sum() == 1
    sph.Run()                                                                    # neighbours get flag 2 (highlight.glsl)
    _, _, hl = sph.download(highlight=True)
    assert hl[77] == 1 and (hl == 2).sum() > 10
    sph.toggle_highlight(77)
    sph.Run()
    _, _, hl = sph.download(highlight=True)
    assert not hl.any()
    with pytest.raises(RuntimeError):
        sph.toggle_highlight(pos.shape[0])


def test_option_density_self_term(built_lib):
    """pbf_options::density_self_term adds W(0) back to rho_i (the reference skips j == i): lambda scales by C'/C."""
    pos, vel = oracle.dam_break(32, 32, 32)
    g = oracle.make_grid(*GRID)
    sph = pbf_b200.SPH(pos.shape[0], GRID)
    sph.upload(pos, vel)
    P = oracle_params(sph)
    sph.predict(); sph.sort(); sph.build_cells()
    _, _, rec = sph.get_sorted()
    start, end = oracle.findcells(rec, g)
    rs, rc = oracle.neighbourcells(rec, g, start, end)
    olam, orho = oracle.calclambda(rec, rs, rc, P)
    sph.calc_lambda()
    assert np.max(np.abs(sph.get_lambda() - olam)) < 1e-5 * np.max(np.abs(olam))      # default: the reference's lambda
    sph.set_options(density_self_term=True)
    assert sph.get_options().density_self_term == 1
    sph.calc_lambda()
    lam = sph.get_lambda()
    w0 = np.float32(pbf_b200.wpoly6(0.0, 2.0))
    c_ref = orho.astype(np.float64) * P.one_over_rho_0 - 1.0
    c_new = (orho.astype(np.float64) + w0) * P.one_over_rho_0 - 1.0
    ok = np.abs(c_ref) > 1e-3                                                         # lambda = -C / (S + eps): same S
    want = olam.astype(np.float64)[ok] * c_new[ok] / c_ref[ok]
    assert np.max(np.abs(lam[ok] - want)) < 2e-4 * np.max(np.abs(want))
    sph.set_options(density_self_term=False)
    sph.calc_lambda()
    assert np.max(np.abs(sph.get_lambda() - olam)) < 1e-5 * np.max(np.abs(olam))


def test_option_wall_restitution(built_lib):
    """Non-interacting particles dropped onto the floor: the reference leaves them the velocity the clamp implies
    (downwards); with wall_restitution e they leave the floor with -e times that."""
    n = 512
    pos = np.zeros((n, 4), np.float32)
    k = np.arange(n)
    pos[:, 0] = 20.0 + 3.0 * (k % 16); pos[:, 2] = 20.0 + 2.8 * (k // 16); pos[:, 1] = 0.05      # > h apart: no neighbours
    vel = np.zeros((n, 4), np.float32)
    vel[:, 1] = -10.0
    out = {}
    for e in (None, 0.0, 0.5):
        sph = pbf_b200.SPH(n, GRID)
        sph.SetNumSolverIterations(2)
        if e is not None:
            sph.set_options(wall_restitution=e)
        sph.upload(pos, vel)
        sph.Run()
        out[e] = sph.download()
    p, v = out[None]
    assert np.all(p[:, 1] == 0.0) and np.allclose(v[:, 1], -0.05 / 0.016, rtol=1e-5)     # update.glsl: (0 - 0.05) / dt
    assert np.all(out[0.0][0][:, 1] == 0.0) and np.all(out[0.0][1][:, 1] == 0.0)
    assert np.allclose(out[0.5][1][:, 1], 0.5 * 0.05 / 0.016, rtol=1e-5)
    assert np.array_equal(out[0.5][1][:, [0, 2]], v[:, [0, 2]])                          # tangential components untouched
    with pytest.raises(RuntimeError):
        pbf_b200.SPH(512, (16, 16, 16)).set_options(wall_restitution=1.5)


@pytest.mark.parametrize("seed", range(10))
def test_randomized_scenes(built_lib, seed):
    """Seeded random grids, walls, parameters and particle mixtures (a lattice blob, a uniform scatter, a few particles
    outside the grid, a few exact duplicates): integer tables bit exact, two whole steps within tolerance."""
    rng = np.random.default_rng(1000 + seed)
    grid = (int(rng.integers(24, 97)), int(rng.integers(12, 49)), int(rng.integers(24, 97)))
    wall = (float(rng.integers(0, 5)), 0.0, float(rng.integers(0, 5)))
    quirks = bool(rng.integers(0, 2))
    n = 512 * int(rng.integers(1, 7))
    nb = n // 2
    side = int(np.ceil(nb ** (1 / 3)))
    ii = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 3)[:nb]
    lo = np.array([wall[0] + 1, 0.5, wall[2] + 1])
    hi = np.array([grid[0] - wall[0] - 1, grid[1] - 1, grid[2] - wall[2] - 1], np.float64)
    spacing = float(rng.uniform(0.7, 1.1))
    blob = np.minimum(lo + spacing * ii + rng.uniform(-0.01, 0.01, (nb, 3)), hi)
    scatter = rng.uniform(lo, hi, (n - nb, 3))
    xyz = np.concatenate([blob, scatter])
    out = rng.choice(n, 8, replace=False)
    xyz[out] += rng.choice([-1.0, 1.0], (8, 3)) * np.array(grid) * (rng.random((8, 3)) < 0.4)     # some far outside
    dup = rng.choice(n, 6, replace=False)
    xyz[dup[:3]] = xyz[dup[3:]]                                                                    # exact duplicates
    pos = np.zeros((n, 4), np.float32); pos[:, :3] = xyz
    vel = np.zeros((n, 4), np.float32); vel[:, :3] = rng.normal(0, float(rng.uniform(0, 6)), (n, 3))
    iters = int(rng.integers(1, 4))
    vort = bool(rng.integers(0, 2))
    g = oracle.make_grid(*grid, wall=wall, ref_quirks=int(quirks))
    sph = pbf_b200.SPH(n, grid, wall=wall, ref_quirks=quirks)
    sph.SetNumSolverIterations(iters)
    sph.SetVorticityConfinementEnabled(vort)
    sph.SetGravity(float(rng.uniform(5, 15)))
    sph.SetTimestep(float(rng.uniform(0.008, 0.02)))
    sph.SetCFMEpsilon(float(rng.uniform(1, 10)))
    sph.upload(pos, vel)
    check_tables(sph, g, pos, vel, quirks)
    sph.SetExternalForce(bool(rng.integers(0, 2)))       # after the table check, which predicts without it
    P = oracle_params(sph)
    extforce = bool(sph._get().external_force)
    dt = sph.GetTimestep()
    sph.upload(pos, vel)
    sim = oracle.Sim(n, g)
    opos, ovel = pos.copy(), vel.copy()
    for step in range(2):
        before = opos.copy()
        sph.Run()
        sim.step(opos, ovel, P, iters, vorticity=vort, extforce=extforce)
        gpos, gvel = sph.download()
        scale = max(1.0, np.max(np.abs(opos - before)) / 0.5)
        assert np.max(np.abs(gpos - opos)) < POS_TOL * scale, (seed, step)
        assert np.max(np.abs(gvel - ovel)) < POS_TOL * scale / dt, (seed, step)
        sph.upload(opos, ovel)
