#!/usr/bin/env python
"""Times profiles/stage_times.py once per library variant under pbf_b200/variants/ (and the product build first)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
libs = [None] + sorted(glob.glob(os.path.join(ROOT, "pbf_b200", "variants", "libpbf_b200_*.so")))
for lib in libs:
    env = dict(os.environ)
    if lib:
        env["PBF_B200_LIB"] = lib
    print("==== %s" % (os.path.basename(lib) if lib else "product build"), flush=True)
    subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "stage_times.py")] + sys.argv[1:], env=env)
