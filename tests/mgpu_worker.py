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
    s.upload_slab(p, v, g)
    steps = 6
    s.Run(steps)
    lp, lv, lg = s.download_slab()
    st = s.stats()
    parts = [None] * world
    dist.all_gather_object(parts, (lp, lv, lg, st))
    ok = True
    if rank == 0:
        gpos = np.zeros_like(pos); gvel = np.zeros_like(vel); seen = np.zeros(pos.shape[0], np.int32)
        for lp, lv, lg, _ in parts:
            gpos[lg], gvel[lg] = lp, lv
            seen[lg] += 1
        single = pbf_b200.SPH(pos.shape[0], grid, ref_quirks=False, device=local)
        single.SetNumSolverIterations(3)
        single.SetVorticityConfinementEnabled(True)
        single.upload(pos, vel)
        single.Run(steps)
        spos, svel = single.download()
        dp, dv = np.max(np.abs(spos - gpos)), np.max(np.abs(svel - gvel))
        mig = sum(p[3]["migrated"] for p in parts)
        gh = sum(p[3]["ghosts_lo"] + p[3]["ghosts_hi"] for p in parts)
        ok = bool(np.all(seen == 1) and dp < 2e-4 and dv < 2e-4 / 0.016 and mig > 0 and gh > 0)
        print("MGPU_RESULT ok=%s p2p=%s world=%d dp=%.3g dv=%.3g migrated=%d ghosts=%d planes=%s" % (ok, p2p, world, dp, dv, mig, gh, planes))
    dist.barrier()
    s.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
