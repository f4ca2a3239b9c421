#!/usr/bin/env python
"""Mints tests/golden/ref_*.npz from the REFERENCE ITSELF: the shaders of /root/reference/shaders compiled verbatim by g++
(oracle/_ref/libpbf_ref.so, see oracle/ref_harness.cpp), run through SPH::Run's own dispatch sequence.

    python tests/golden/make_ref_golden.py        (only where /root/reference exists; the .npz files are committed)

ref_c1_trace.npz   BASELINE configs[0]: dam-break 32^3 = 32,768 particles, grid 128x64x128, 3 solver iterations, vorticity
                   off, 100 steps.  SHA-256 of the position / velocity bits after steps 1, 10, 100 and per-step kinetic
                   energy (from the reference's velocities) -- the C oracle must reproduce the digests bit for bit.
ref_small.npz      16^3 = 4,096 particles, same grid, K = 3, vorticity + XSPH on: full state after steps 1, 10 and 100,
                   sorted records / cell starts / packed neighbour runs / lambda of step 1, per-step kinetic energy.
ref_reference_scene.npz  the reference's own scene (two mirrored 32^3 blocks = 65,536 particles, K = 5, its defaults,
                   src/Simulation.cpp:200-246 with a seeded jitter): digests after steps 1 and 5.
ref_c2_step.npz    BASELINE configs[1]: dam-break 128x64x128 = 1,048,576 particles, grid 256x128x256, K = 3, vorticity +
                   XSPH on: digests of positions and velocities after steps 1 and 2 (22 s of the reference per step).
ref_c3_step.npz    BASELINE configs[2], the headline size: dam-break 256x128x256 = 8,388,608 particles, grid 512x256x512, K = 4,
                   vorticity + XSPH on: digests after step 1 (about three and a half minutes of the reference).  On this
                   grid (2^26 cells) the shaders' float-dot sort key is inexact: the oracle matches with ref_quirks = 3.
Schedule of the two racy shaders: Jacobi (oracle/ref_harness.cpp, order 0); `define_last_end` policy on."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle
from oracle import ref

HERE = os.path.dirname(os.path.abspath(__file__))
GRID = (128, 64, 128)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def kinetic(vel):
    v = vel[:, :3].astype(np.float64)
    return float(0.5 * np.sum(v * v))


def c1_trace():
    pos, vel = oracle.dam_break(32, 32, 32)
    r = ref.RefSim(pos.shape[0], GRID)
    r.upload(pos, vel)
    out = {"n3": np.array([32, 32, 32]), "grid": np.array(GRID), "iters": 3, "vorticity": 0, "seed": 12345, "steps": 100}
    ke = []
    for step in range(1, 101):
        r.step(3, vorticity=False)
        p, v, _ = r.download()
        ke.append(kinetic(v))
        if step in (1, 10, 100):
            out["pos_sha_%d" % step], out["vel_sha_%d" % step] = digest(p), digest(v)
    out["kinetic_energy"] = np.array(ke)
    np.savez_compressed(os.path.join(HERE, "ref_c1_trace.npz"), **out)


def small():
    pos, vel = oracle.dam_break(16, 16, 16)
    r = ref.RefSim(pos.shape[0], GRID)
    r.upload(pos, vel)
    out = {"n3": np.array([16, 16, 16]), "grid": np.array(GRID), "iters": 3, "vorticity": 1, "seed": 12345, "steps": 100,
           "pos0": pos}
    ke = []
    for step in range(1, 101):
        if step == 1:                      # stage by stage, to keep the intermediate tables of the first step
            r.predict(); r.sort(); r.find_cells(); r.policy_define_last_end(); r.neighbour_cells()
            out["sorted1"] = r.records()
            out["start1"] = r.grid_tables()[0]
            out["runs1"] = r.packed_runs()[0]
            r.highlight()
            for it in range(3):
                r.calclambda()
                if it == 0:
                    out["lambda1"] = r.lam()
                r.updatepos()
            r.update(); r.vorticity()
        else:
            r.step(3, vorticity=True)
        p, v, _ = r.download()
        ke.append(kinetic(v))
        if step in (1, 10, 100):
            out["pos%d" % step], out["vel%d" % step] = p, v
    out["kinetic_energy"] = np.array(ke)
    np.savez_compressed(os.path.join(HERE, "ref_small.npz"), **out)


def reference_scene():
    p1, v1 = oracle.dam_break(32, 32, 32)
    p2, v2 = oracle.dam_break(32, 32, 32, origin=(32.5 + 63.0, 0.5, 32.5 + 63.0), mirror=True, id0=32768)
    pos, vel = np.concatenate([p1, p2]), np.concatenate([v1, v2])
    r = ref.RefSim(pos.shape[0], GRID)
    r.upload(pos, vel)
    out = {"iters": 5, "vorticity": 0, "seed": 12345}
    for step in range(1, 6):
        r.step(5, vorticity=False)
        if step in (1, 5):
            p, v, _ = r.download()
            out["pos_sha_%d" % step], out["vel_sha_%d" % step] = digest(p), digest(v)
    np.savez_compressed(os.path.join(HERE, "ref_reference_scene.npz"), **out)


def c2_step():
    grid = (256, 128, 256)
    pos, vel = oracle.dam_break(128, 64, 128)
    r = ref.RefSim(pos.shape[0], grid)
    r.upload(pos, vel)
    out = {"n3": np.array([128, 64, 128]), "grid": np.array(grid), "iters": 3, "vorticity": 1, "seed": 12345}
    for step in (1, 2):
        r.step(3, vorticity=True)
        p, v, _ = r.download()
        out["pos_sha_%d" % step], out["vel_sha_%d" % step] = digest(p), digest(v)
    np.savez_compressed(os.path.join(HERE, "ref_c2_step.npz"), **out)


def c3_step():
    grid = (512, 256, 512)
    pos, vel = oracle.dam_break(256, 128, 256)
    r = ref.RefSim(pos.shape[0], grid)
    r.upload(pos, vel)
    out = {"n3": np.array([256, 128, 256]), "grid": np.array(grid), "iters": 4, "vorticity": 1, "seed": 12345}
    r.step(4, vorticity=True)
    p, v, _ = r.download()
    out["pos_sha_1"], out["vel_sha_1"] = digest(p), digest(v)
    np.savez_compressed(os.path.join(HERE, "ref_c3_step.npz"), **out)


if __name__ == "__main__":
    if not ref.available():
        raise SystemExit("needs /root/reference (the shaders are compiled from there)")
    only = sys.argv[1:]
    for name, fn in (("c1_trace", c1_trace), ("small", small), ("reference_scene", reference_scene), ("c2_step", c2_step),
                     ("c3_step", c3_step)):
        if not only or name in only:
            fn()
    for f in ("ref_c1_trace.npz", "ref_small.npz", "ref_reference_scene.npz", "ref_c2_step.npz", "ref_c3_step.npz"):
        if not os.path.exists(os.path.join(HERE, f)):
            continue
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
