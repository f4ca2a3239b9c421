"""State files (dump / resume): the host-side format on CPU, the bit-exact resume on the GPU."""
import os

import numpy as np
import pytest

import pbf_b200


def scene():
    pos, vel = pbf_b200.dam_break(16, 8, 8, seed=3)
    vel[:, :3] = np.random.default_rng(0).normal(0, 2, (pos.shape[0], 3)).astype(np.float32)
    hl = np.zeros(pos.shape[0], np.uint32)
    hl[[1, 17, 500]] = [1, 2, 3]
    return pos, vel, hl


def test_round_trip_on_host(built_lib, tmp_path):
    pos, vel, hl = scene()
    p = pbf_b200.default_params()
    p.num_solver_iterations = 3
    p.vorticity_confinement = 1
    p.gravity = 7.5
    path = tmp_path / "a.pbfstate"
    pbf_b200.write_state_file(path, pos, vel, hl, grid=(100, 50, 90), wall=(8.0, 0.0, 4.0), ref_quirks=False, params=p, steps=41,
                              options=pbf_b200.Options(1, 0.25))
    assert os.path.getsize(path) == 128 + pos.shape[0] * 36
    assert not os.path.exists(str(path) + ".part")
    info, rpos, rvel, rhl = pbf_b200.read_state_file(path)
    assert info.num_particles == pos.shape[0] and tuple(info.grid) == (100, 50, 90) and tuple(info.wall) == (8.0, 0.0, 4.0)
    assert info.ref_quirks == 0 and info.steps == 41
    assert info.options.density_self_term == 1 and info.options.wall_restitution == 0.25
    assert info.params.num_solver_iterations == 3 and info.params.vorticity_confinement == 1 and info.params.gravity == 7.5
    assert np.array_equal(rpos.view(np.uint32), pos.view(np.uint32))
    assert np.array_equal(rvel.view(np.uint32), vel.view(np.uint32))
    assert np.array_equal(rhl, hl)
    # header layout is part of the contract: magic, version, header size, N
    raw = open(path, "rb").read(128)
    assert raw[:8] == b"PBFB200S"
    assert np.frombuffer(raw[8:16], np.uint32).tolist() == [1, 128]
    assert int(np.frombuffer(raw[16:24], np.uint64)[0]) == pos.shape[0]


def test_missing_arrays_are_zero(built_lib, tmp_path):
    pos, _, _ = scene()
    path = tmp_path / "b.pbfstate"
    pbf_b200.write_state_file(path, pos)
    info, rpos, rvel, rhl = pbf_b200.read_state_file(path)
    assert np.array_equal(rpos, pos) and not rvel.any() and not rhl.any()
    assert info.options.density_self_term == 0 and info.options.wall_restitution < 0      # corrections off by default


def test_corruption_is_detected(built_lib, tmp_path):
    pos, vel, hl = scene()
    path = tmp_path / "c.pbfstate"
    pbf_b200.write_state_file(path, pos, vel, hl)
    raw = bytearray(open(path, "rb").read())
    bad = bytearray(raw); bad[128 + 1000] ^= 0x40                 # one flipped payload bit
    open(tmp_path / "flip", "wb").write(bad)
    with pytest.raises(RuntimeError, match="checksum"):
        pbf_b200.read_state_file(tmp_path / "flip")
    open(tmp_path / "short", "wb").write(raw[:-100])
    with pytest.raises(RuntimeError, match="truncated"):
        pbf_b200.read_state_file(tmp_path / "short")
    bad = bytearray(raw); bad[0] = ord("X")
    open(tmp_path / "magic", "wb").write(bad)
    with pytest.raises(RuntimeError, match="not a pbf_b200 state file"):
        pbf_b200.read_state_file(tmp_path / "magic")
    bad = bytearray(raw); bad[8] = 9
    open(tmp_path / "version", "wb").write(bad)
    with pytest.raises(RuntimeError, match="version"):
        pbf_b200.state_file_info(tmp_path / "version")
    with pytest.raises(RuntimeError, match="cannot open"):
        pbf_b200.state_file_info(tmp_path / "nope")


@pytest.mark.gpu
@pytest.mark.parametrize("graph", [False, True])
def test_resume_is_bit_exact(built_lib, tmp_path, graph):
    """10 steps == 4 steps, save, new handle from the file, 6 steps -- bit for bit (the step is a pure function of
    positions, velocities, highlight flags and parameters)."""
    pos, vel = pbf_b200.dam_break(32, 32, 32)
    a = pbf_b200.SPH(pos.shape[0], use_graph=graph)
    a.SetNumSolverIterations(3)
    a.SetVorticityConfinementEnabled(True)
    a.upload(pos, vel)
    a.Run(4)
    path = tmp_path / "mid.pbfstate"
    a.save_state(path)
    a.Run(6)
    apos, avel = a.download()
    b = pbf_b200.SPH.from_state_file(path, use_graph=graph)
    assert b.step_count == 4 and b.GetNumSolverIterations() == 3 and b.IsVorticityConfinementEnabled()
    b.Run(6)
    assert b.step_count == 10 == a.step_count
    bpos, bvel = b.download()
    assert np.array_equal(apos.view(np.uint32), bpos.view(np.uint32))
    assert np.array_equal(avel.view(np.uint32), bvel.view(np.uint32))
    # a handle of another size refuses the file
    c = pbf_b200.SPH(512, (16, 16, 16))
    with pytest.raises(RuntimeError, match="differs"):
        c.load_state(path)


@pytest.mark.gpu
def test_gl_registration_fails_cleanly_without_a_context(built_lib):
    """No GL context on the GPU box: registration must fail with a CUDA error, leave the handle usable and on its own
    buffers (the success path needs the renderer and cannot run here)."""
    pos, vel = pbf_b200.dam_break(16, 16, 16)
    sph = pbf_b200.SPH(pos.shape[0])
    with pytest.raises(RuntimeError, match="cudaGraphicsGLRegisterBuffer"):
        sph.register_gl_buffers(1, 2, 3)
    sph.unregister_gl_buffers()          # no-op
    sph.upload(pos, vel)
    sph.Run(2)
    p, _ = sph.download()
    assert np.isfinite(p).all() and not np.array_equal(p, pos)
