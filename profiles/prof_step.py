#!/usr/bin/env python
"""Profiling driver: the C3 headline scene, a few warm steps through the graph, then ONE step through the stage entry
points (direct launches, so `ncu -k regex:...` sees plain kernels).

  ncu --set full --clock-control none --import-source on -k regex:'k_lambda|k_delta_p|k_vorticity|k_plan' \
      --launch-skip 0 --launch-count 8 -o gpurun_out/prof_X python profiles/prof_step.py
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pbf_b200

small = "--small" in sys.argv
n3, grid = ((128, 64, 128), (256, 128, 256)) if small else ((256, 128, 256), (512, 256, 512))
pos, vel = pbf_b200.dam_break(*n3)
sph = pbf_b200.SPH(pos.shape[0], grid, ref_quirks=False, use_graph=False)      # direct launches: ncu -k sees plain kernels
sph.SetNumSolverIterations(4)
sph.SetVorticityConfinementEnabled(True)
sph.upload(pos, vel)
sph.Run(int(os.environ.get("PROF_WARM", "3")))
sph.sync()
sph.predict(); sph.sort(); sph.build_cells()
print("tiles, tiled:", sph.tile_stats(), sph.tile_fallback_reasons)
for _ in range(2):
    sph.calc_lambda(); sph.update_positions()
sph.finalize(); sph.vorticity()
sph.sync()
sph.Run(1)                       # one whole step: the fused kernels (predict + table reset, last delta-p + update)
sph.sync()
print("ok")
