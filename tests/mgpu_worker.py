"""Worker for test_multi_gpu.py: run under torchrun, one rank per GPU.  Every rank builds the same seeded scene, owns
one z-slab of it (NCCL halo exchange + migration inside libpbf_b200), and rank 0 compares the gathered result with a
single-domain run of the same scene on its own GPU."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pbf_b200
from pbf_b200 import slab


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    grid = (64, 32, 160)
    pos, vel = pbf_b200.dam_break(16, 16, 128, origin=(18.5, 0.5, 18.5))
    vel[:, :3] = np.random.default_rng(11).normal(0, 4.0, (pos.shape[0], 3)).astype(np.float32)
    planes = slab.plan_slabs(slab.cell_layer(pos, grid[2]), grid[2], world)
    p, v, g = slab.split_scene(pos, vel, planes, grid[2])[rank]
    s = slab.SlabSPH(rank, world, planes, grid[:2], grid[2], int(p.shape[0] * 1.5) + 2 * 8192, 8192, device=local)
    s.init_nccl(slab.broadcast_unique_id(dist, rank, torch.device("cuda", local)))
    p2p = s.connect_p2p(dist, torch.device("cuda", local))     # PBF_SLAB_P2P=0: NCCL send/recv for every exchange
    s.SetNumSolverIterations(3)
    s.SetVorticityConfinementEnabled(True)
    s.SetExternalForce(True)            # predictpos.glsl:27 compares against GRID_SIZE.z/2 of the WHOLE domain (80 here)
    canonical = os.environ.get("PBF_TEST_CANONICAL") == "1"      # one summation order everywhere: results must be BIT exact
    s.set_canonical_order(canonical)
    s.upload_slab(p, v, g)
    # select the particles of one column either side of the first plane: highlight.glsl's marks must cross it
    sel = np.nonzero((np.abs(pos[:, 2] - planes[1]) < 1.0) & (np.abs(pos[:, 0] - 25.0) < 1.0) & (pos[:, 1] < 3.0))[0]
    for slot in np.nonzero(np.isin(g, sel))[0]:
        s.toggle_highlight(int(slot))
    steps = 6
    s.Run(steps)
    lp, lv, lg = s.download_slab()
    lh = s.download_highlight()
    st = s.stats()
    parts = [None] * world
    dist.all_gather_object(parts, (lp, lv, lg, st, lh))
    ok = True
    if rank == 0:
        gpos = np.zeros_like(pos); gvel = np.zeros_like(vel); seen = np.zeros(pos.shape[0], np.int32)
        ghl = np.zeros(pos.shape[0], np.uint32)
        for lp, lv, lg, _, lh in parts:
            gpos[lg], gvel[lg], ghl[lg] = lp, lv, lh
            seen[lg] += 1
        single = pbf_b200.SPH(pos.shape[0], grid, ref_quirks=False, device=local)
        single.SetNumSolverIterations(3)
        single.SetVorticityConfinementEnabled(True)
        single.SetExternalForce(True)
        single.set_canonical_order(canonical)
        single.upload(pos, vel)
        for i in sel:
            single.toggle_highlight(int(i))
        single.Run(steps)
        spos, svel, shl = single.download(highlight=True)
        dp, dv = np.max(np.abs(spos - gpos)), np.max(np.abs(svel - gvel))
        mig = sum(p[3]["migrated"] for p in parts)
        gh = sum(p[3]["ghosts_lo"] + p[3]["ghosts_hi"] for p in parts)
        # selection bits are bit exact; the marks of the last step depend on which particles are neighbours, and a particle
        # within rounding distance of a cell face may sit on either side of it in the two runs: allow a handful
        hl_bad = int(np.count_nonzero(shl != ghl))
        marked = int(np.count_nonzero(ghl & 2))
        ok = bool(np.all(seen == 1) and dp < 2e-4 and dv < 2e-4 / 0.016 and mig > 0 and gh > 0
                  and np.array_equal(shl & 1, ghl & 1) and sel.size >= 4 and marked > sel.size and hl_bad <= 2)
        if canonical:
            ok = ok and np.array_equal(spos.view(np.uint32), gpos.view(np.uint32)) and np.array_equal(svel.view(np.uint32), gvel.view(np.uint32)) \
                and hl_bad == 0
        print("MGPU_RESULT ok=%s canonical=%s p2p=%s world=%d dp=%.3g dv=%.3g migrated=%d ghosts=%d selected=%d marked=%d hl_mismatch=%d planes=%s"
              % (ok, canonical, p2p, world, dp, dv, mig, gh, sel.size, marked, hl_bad, planes))
    dist.barrier()
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
