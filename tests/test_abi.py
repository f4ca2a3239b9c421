"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
include/pbf_c.h declares.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "pbf_c.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pbf_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_host_only_entry_points(built_lib):
    import pbf_b200
    import oracle
    assert abs(pbf_b200.wpoly6(0.0, 2.0) - 0.19583518) < 1e-7
    p = pbf_b200.default_params()
    assert p.num_solver_iterations == 5 and p.vorticity_confinement == 0      # src/SPH.cpp:25-26
    o = oracle.default_params()
    for k in ("one_over_rho_0", "epsilon", "gravity", "timestep", "tensile_instability_k",
              "tensile_instability_scale", "xsph_viscosity_c", "vorticity_epsilon"):
        assert getattr(p, k) == getattr(o, k)
    assert pbf_b200.sort_bits((128, 64, 128)) == 20       # 10 two-bit passes (src/RadixSort.cpp:127)
    assert pbf_b200.sort_bits((256, 128, 256)) == 24
    assert pbf_b200.sort_bits((512, 256, 512)) == 26
    # onesweep passes of up to 9 bits over the key bits that can be set at all: no hash exceeds ncell + gx*gz + gx, so a slab
    # rank's 514-layer window (28 bits by the reference's even-bits rule, 27 live) sorts in three passes like the 512-layer grid
    assert pbf_b200.sort_passes((512, 256, 512)) == 3
    assert pbf_b200.sort_passes((512, 256, 514)) == 3
    assert pbf_b200.sort_passes((1024, 512, 1024)) == 4      # 30 bits
    assert pbf_b200.sort_passes((128, 64, 128)) == 3         # 20 bits
    assert pbf_b200.sort_passes((64, 32, 64)) == 2           # 18 bits (the clamped-cell hash may reach 2^17 + ...: still 18)
    a, av = pbf_b200.dam_break(8, 4, 6, seed=99)
    b, bv = oracle.dam_break(8, 4, 6, seed=99)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)) and not av.any()


def test_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import pbf_b200
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        pbf_b200.SPH(512)


def test_header_is_plain_c():
    """include/pbf_c.h is a C header (no C++ or torch types in the signatures): it must compile as C99 with warnings on."""
    import subprocess
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c",
                        os.path.join(ROOT, "include", "pbf_c.h")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
