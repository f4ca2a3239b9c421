#!/usr/bin/env python
"""How far the reference's literal sort key (a float dot product, exact below 2^24) takes it from its own algorithm on the
headline grid of BASELINE configs[2] (512x256x512 = 2^26 cells): one step of the oracle with the literal key (ref_quirks = 3,
bit-identical to the compiled shaders: tests/test_oracle_ref.py) against one step with the integer hash (ref_quirks = 1, what
the product implements).  Prints the numbers quoted in DESIGN.md section 2.

    python tests/golden/float_hash_departure.py            (CPU, about half a minute)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle

n3, grid = (256, 128, 256), (512, 256, 512)
out = {}
for q in (1, 3):
    pos, vel = oracle.dam_break(*n3)
    sim = oracle.Sim(pos.shape[0], oracle.make_grid(*grid, ref_quirks=q))
    sim.step(pos, vel, oracle.default_params(), 4, vorticity=True)
    out[q] = (pos.copy(), vel.copy())
dp = np.abs(out[1][0][:, :3] - out[3][0][:, :3]).max(axis=1)
dv = np.abs(out[1][1][:, :3] - out[3][1][:, :3]).max(axis=1)
tol = 1e-5 * 128
print("particles: %d, differing at all: %d (%.2f %%), beyond the one-step tolerance %.3g: %d (%.2f %%)"
      % (dp.size, int((dp > 0).sum()), 100.0 * (dp > 0).mean(), tol, int((dp > tol).sum()), 100.0 * (dp > tol).mean()))
print("max |dpos| = %.4g cells, max |dvel| = %.4g" % (dp.max(), dv.max()))
