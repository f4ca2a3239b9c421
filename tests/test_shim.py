"""The C++ shim classes (include/pbf/*.h: SPH, RadixSort, NeighbourCellFinder, Simulation) compile and link against
libpbf_b200.so (CPU check) and, on the GPU box, produce the same state as the Python binding of the same C ABI."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "pbf_b200", "build", "shim_smoke")


def build_exe(lib):
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "shim_smoke.cpp"), "-o", EXE, lib, "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_shim_compiles_and_links(built_lib):
    build_exe(built_lib)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_shim_matches_python_binding(built_lib, tmp_path):
    import pbf_b200
    exe = build_exe(built_lib)
    dump = str(tmp_path / "shim_state.bin")
    r = subprocess.run([exe, "3", dump], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    line = [l for l in r.stdout.splitlines() if l.startswith("SHIM n=")][0]
    vals = dict(kv.split("=") for kv in line.split()[1:])
    assert "expected error" in r.stdout
    assert "SHIM pick=1 again=1" in r.stdout           # Simulation::OnMouseDown: the ray finds a particle, twice the same
    p1, v1 = pbf_b200.dam_break(32, 32, 32)
    p2, v2 = pbf_b200.dam_break(32, 32, 32, origin=(95.5, 0.5, 95.5), mirror=True, id0=32768)
    sph = pbf_b200.SPH(65536)
    sph.SetNumSolverIterations(3)
    sph.SetVorticityConfinementEnabled(True)
    sph.upload(np.concatenate([p1, p2]), np.concatenate([v1, v2]))
    sph.Run(4)
    pos, vel = sph.download()
    assert abs(float(vals["sum_pos"]) - pos.astype(np.float64).sum()) < 1e-3 * 65536
    assert abs(float(vals["sum_v2"]) - (vel.astype(np.float64) ** 2).sum()) < 1e-3 * float(vals["sum_v2"])
    # the state itself, particle by particle: three frames through Simulation::Frame (pbf_step) plus one step through the
    # RadixSort / NeighbourCellFinder shims (the stage entry points) against four pbf_step calls of the Python binding --
    # the same kernels on the same inputs, so the same bits
    raw = np.fromfile(dump, np.float32)
    assert raw.size == 2 * 4 * 65536
    spos, svel = raw[:4 * 65536].reshape(-1, 4), raw[4 * 65536:].reshape(-1, 4)
    assert np.array_equal(spos.view(np.uint32), pos.view(np.uint32))
    assert np.array_equal(svel.view(np.uint32), vel.view(np.uint32))
