#!/usr/bin/env python
"""Experiment aid: the lambda kernel alone on the untouched C3 lattice (predict, sort, cells, then k_lambda x reps), once
per library variant under pbf_b200/variants/ -- probes that compute wrong results never get to step the scene."""
import glob
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
if "--child" in sys.argv:
    sys.path.insert(0, ROOT)
    import torch
    import pbf_b200
    pos, vel = pbf_b200.dam_break(256, 128, 256)
    sph = pbf_b200.SPH(pos.shape[0], (512, 256, 512), ref_quirks=False)
    sph.SetNumSolverIterations(4)
    sph.upload(pos, vel)
    sph.predict(); sph.sort(); sph.build_cells()
    stream = torch.cuda.ExternalStream(sph.stream)
    with torch.cuda.stream(stream):
        for name, fn in (("lambda", sph.calc_lambda),):
            fn(); fn()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
            torch.cuda.synchronize()
            e[0].record()
            for k in range(10):
                fn()
                e[k + 1].record()
            torch.cuda.synchronize()
            t = sorted(e[k].elapsed_time(e[k + 1]) for k in range(10))
            print("%s ms: median %.4f min %.4f" % (name, t[5], t[0]))
    sys.exit(0)
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "pbf_b200", "variants", "libpbf_b200_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["PBF_B200_LIB"] = lib
    print("==== %s" % (os.path.basename(lib) if lib else "product build"), flush=True)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env)
