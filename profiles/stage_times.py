#!/usr/bin/env python
"""Per-stage CUDA-event times of the C3 headline scene through the stage entry points (quick iteration aid).

  python profiles/stage_times.py [--small] [--warm N]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import pbf_b200

small = "--small" in sys.argv
warm = int(sys.argv[sys.argv.index("--warm") + 1]) if "--warm" in sys.argv else 5
n3, grid = ((128, 64, 128), (256, 128, 256)) if small else ((256, 128, 256), (512, 256, 512))
pos, vel = pbf_b200.dam_break(*n3)
sph = pbf_b200.SPH(pos.shape[0], grid, ref_quirks=False)
sph.SetNumSolverIterations(4)
sph.SetVorticityConfinementEnabled(True)
sph.upload(pos, vel)
stream = torch.cuda.ExternalStream(sph.stream)


def timed(fn, reps=5):
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps


sph.Run(warm)
sph.sync()
print("step ms (graph): %.3f" % timed(lambda: sph.Run(1), 20))
sph.predict(); sph.sort()
print("build_cells ms: %.3f" % timed(sph.build_cells, 3))
print("tiles, tiled:", sph.tile_stats(), sph.tile_fallback_reasons)
print("lambda ms: %.3f" % timed(sph.calc_lambda))
print("delta_p ms: %.3f" % timed(sph.update_positions))
sph.calc_lambda(); sph.finalize()
print("vorticity a+b ms: %.3f" % timed(sph.vorticity))
sph.upload(pos, vel)
sph.enable_timing(True)
sph.Run(3)
print("phases:", dict(zip(["predict", "sort", "cells", "solver", "vorticity"], ["%.3f" % x for x in sph.get_timings()])))
