#!/usr/bin/env python
"""Headless dam break through the Python binding -- the reference's default scene (two mirrored 32^3 blocks,
src/Simulation.cpp:206-246) or a larger lattice -- with per-phase timings, diagnostics and a state file at the end.

    python examples/run_dam_break.py [--n3 64 32 64] [--grid 128 64 128] [--steps 200] [--iters 3] [--vorticity]
                                     [--save out.pbfstate] [--resume in.pbfstate]

Needs a B200 (the library has no CPU path)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import pbf_b200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n3", type=int, nargs=3, default=None, help="lattice block; default: the reference's two 32^3 blocks")
    ap.add_argument("--grid", type=int, nargs=3, default=(128, 64, 128))
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--vorticity", action="store_true")
    ap.add_argument("--save", default=None)
    ap.add_argument("--resume", default=None)
    args = ap.parse_args()

    if args.resume:
        sph = pbf_b200.SPH.from_state_file(args.resume)
        print("resumed %d particles at step %d from %s" % (sph.numparticles, sph.step_count, args.resume))
    else:
        if args.n3:
            pos, vel = pbf_b200.dam_break(*args.n3)
        else:
            p1, v1 = pbf_b200.dam_break(32, 32, 32)
            p2, v2 = pbf_b200.dam_break(32, 32, 32, origin=(95.5, 0.5, 95.5), mirror=True, id0=32768)
            pos, vel = np.concatenate([p1, p2]), np.concatenate([v1, v2])
        sph = pbf_b200.SPH(pos.shape[0], tuple(args.grid), ref_quirks=False)
        sph.SetNumSolverIterations(args.iters)
        sph.SetVorticityConfinementEnabled(args.vorticity)
        sph.upload(pos, vel)
    report = max(1, args.steps // 10)
    for first in range(0, args.steps, report):
        sph.Run(min(report, args.steps - first))
        err, ke = sph.diagnostics()
        print("step %6d   mean |rho/rho0 - 1| = %.4f   kinetic energy = %.4g" % (sph.step_count, err, ke))
    sph.enable_timing(True)
    sph.Run(1)
    sph.OutputTiming()
    tiles, tiled = sph.tile_stats()
    print("tiles on the shared-memory sweep path: %d of %d" % (tiled, tiles))
    if args.save:
        sph.save_state(args.save)
        print("state written to", args.save)


if __name__ == "__main__":
    main()
