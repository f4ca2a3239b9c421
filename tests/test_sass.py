"""Build-time properties of the hot kernels, checked on the compiled objects without a GPU (cuobjdump ships with CUDA): the
neighbour sweeps stage through TMA bulk copies and prefetches, compute with packed FP32, fit the register budget of eight
blocks per SM and do not spill; the onesweep passes do not spill either.  A refactoring that loses one of these would still
pass every parity test -- and cost tens of per cent."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "pbf_b200", "build")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not installed")


def resources(obj):
    """{mangled kernel name: (registers, stack bytes)}"""
    out = subprocess.run([CUOBJDUMP, "--dump-resource-usage", os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        res[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    return res


def sass_of(obj, pattern):
    out = subprocess.run([CUOBJDUMP, "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True, check=True).stdout
    chunks = re.split(r"\n\s*Function : ", out)
    return [c for c in chunks if re.match(pattern, c)]


def test_sweeps_fit_eight_blocks_per_sm_without_spills(built_lib):
    res = resources("sweeps.o")
    hot = {k: v for k, v in res.items() if re.search(r"k_lambdaILb0E|k_delta_pILi[012]E|k_vorticity_[ab]I", k)}
    assert len(hot) >= 12                                   # every FULL / LOOP instantiation of the four sweeps
    for name, (reg, stack) in hot.items():
        assert stack == 0, (name, stack)                    # no local-memory spills
        assert reg <= 64, (name, reg)                       # 8 blocks x 128 threads (4 x 256 for vorticity A) x 64 registers = one SM


def test_sweeps_use_tma_and_packed_fp32(built_lib):
    # the single-GPU lambda sweep: tiled path, count known on the host
    code = sass_of("sweeps.o", r"\S*k_lambdaILb0ELb0ELb0E")
    assert len(code) == 1
    sass = code[0]
    assert "UBLKCP" in sass                                  # cp.async.bulk: the tile image is staged by the TMA unit
    assert "UBLKPF" in sass                                  # cp.async.bulk.prefetch.L2 one wave ahead
    assert "SYNCS" in sass                                   # mbarrier completion
    for op in ("FFMA2", "FMUL2", "FADD2", "MUFU.RSQ", "LDS.128"):
        assert op in sass, op
    # the pair loop: two 128-bit shared loads and two reciprocal square roots per iteration, nothing from local memory
    assert "LDL" not in sass and "STL" not in sass


def test_onesweep_does_not_spill(built_lib):
    res = resources("sort.o")
    hot = {k: v for k, v in res.items() if "k_onesweep" in k}
    assert hot
    for name, (reg, stack) in hot.items():
        assert stack == 0, (name, stack)
