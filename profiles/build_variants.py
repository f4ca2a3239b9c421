#!/usr/bin/env python
"""Kernel experiments: builds libpbf_b200 variants (sweeps.cu recompiled with extra -D switches, the other objects
reused) into pbf_b200/variants/, to be timed side by side on the GPU box with profiles/time_variants.py.

  python profiles/build_variants.py name1=-DFOO,-DBAR=2 name2=-DBAZ ...
"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from pbf_b200 import build as B

B.build()
out = os.path.join(B.HERE, "variants")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    name, flags = spec.split("=", 1)
    flags = [f for f in flags.split(",") if f]
    objs = []
    for src in B.SOURCES:
        base = src.rsplit(".", 1)[0]
        if src in ("sweeps.cu", "sim_kernels.cu", "sort.cu"):      # the translation units the experiment switches live in
            o = os.path.join(out, "%s_%s.o" % (base, name))
            cmd = [B.NVCC] + B.ARCH + B.CUFLAGS + flags + ["-I", os.path.join(B.HERE, "..", "include"), "-c",
                                                         os.path.join(B.CSRC, src), "-o", o]
            subprocess.run(cmd, check=True)
            objs.append(o)
        else:
            objs.append(os.path.join(B.OBJ, base + ".o"))
    lib = os.path.join(out, "libpbf_b200_%s.so" % name)
    subprocess.run([B.NVCC] + B.ARCH + ["-shared", "-o", lib] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"], check=True)
    print(lib)
