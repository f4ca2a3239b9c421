"""CPU model of the slab algorithm of pbf_b200/csrc/slab.cu, one process per rank over gloo, built from the oracle's
stage functions: predict -> migrate -> ghosts -> sort/cells -> K x (lambda, halo lambda, delta-p, halo positions) ->
update.  Rank 0 checks the gathered result against the single-domain oracle.  It validates the decomposition itself
(one ghost layer suffices, migration after predict, ghosts carry their old position) and the host-side partitioning
and unique-id plumbing of pbf_b200.slab -- no GPU involved."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from pbf_b200 import slab


def exchange(rank, world, to_lo, to_hi):
    """all ranks post what they send down/up; returns (from_lo, from_hi)"""
    box = [None] * world
    dist.all_gather_object(box, (to_lo, to_hi))
    return (box[rank - 1][1] if rank > 0 else None), (box[rank + 1][0] if rank + 1 < world else None)


def cat(*parts):
    parts = [p for p in parts if p is not None and len(p)]
    return np.concatenate(parts) if parts else None


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    grid = (48, 24, 64)
    g = oracle.make_grid(*grid, ref_quirks=0)
    P = oracle.default_params()
    K, steps = 3, 4
    pos, vel = oracle.dam_break(8, 8, 32, origin=(18.5, 0.5, 18.5), seed=77)
    vel[:, :3] = np.random.default_rng(3).normal(0, 5.0, (pos.shape[0], 3)).astype(np.float32)
    planes = slab.plan_slabs(slab.cell_layer(pos, grid[2]), grid[2], world)
    # unique-id plumbing (the id itself needs libnccl + a GPU, so a stand-in is broadcast here)
    slab.unique_id = lambda: bytes(range(128))
    assert slab.broadcast_unique_id(dist, rank) == bytes(range(128))
    lp, lv, lg = [a.copy() for a in slab.split_scene(pos, vel, planes, grid[2])[rank]]
    z_lo, z_hi = planes[rank], planes[rank + 1]
    migrated = 0
    # SLAB_MODEL_CANONICAL=1: the composite (cell key, global id) order of SURVEY.md 8e -- every rank files its local and ghost
    # particles in ascending global id before the stable sort, as the single-domain run does by construction.  The oracle adds a
    # particle's terms in candidate order, so the slab run must then equal the single-domain run BIT FOR BIT.
    canonical = os.environ.get("SLAB_MODEL_CANONICAL") == "1"
    for _ in range(steps):
        n = lp.shape[0]
        rec = oracle.predict(lp, lv, P, g)
        cz = slab.cell_layer(rec, grid[2])
        lo, hi = (cz < z_lo) & (rank > 0), (cz >= z_hi) & (rank + 1 < world)
        pack = lambda m: (lp[m], lv[m], rec[m], lg[m])
        a_lo, a_hi = exchange(rank, world, pack(lo), pack(hi))
        stay = ~(lo | hi)
        migrated += int((~stay).sum())
        arr = [a for a in (a_lo, a_hi) if a is not None]
        lp = cat(lp[stay], *[a[0] for a in arr]); lv = cat(lv[stay], *[a[1] for a in arr])
        rec = cat(rec[stay], *[a[2] for a in arr]); lg = cat(lg[stay], *[a[3] for a in arr])
        n = lp.shape[0]
        cz = slab.cell_layer(rec, grid[2])
        b_lo, b_hi = np.nonzero((cz == z_lo) & (rank > 0))[0], np.nonzero((cz == z_hi - 1) & (rank + 1 < world))[0]
        g_lo, g_hi = exchange(rank, world, (rec[b_lo], lp[b_lo], lg[b_lo]), (rec[b_hi], lp[b_hi], lg[b_hi]))
        ghosts = [x for x in (g_lo, g_hi) if x is not None]
        ng = [x[0].shape[0] for x in ghosts]
        rec_all = cat(rec, *[x[0] for x in ghosts]).copy()
        pos_all = cat(lp, *[x[1] for x in ghosts]).copy()
        vel_all = np.zeros_like(pos_all); vel_all[:n] = lv
        rec_all[:, 3] = np.arange(rec_all.shape[0], dtype=np.int32).view(np.float32)
        if canonical:
            # feed the sort in global-id order: its stability then orders every cell by id (slots keep naming the rows of
            # pos_all / vel_all, so nothing else changes)
            gid_all = cat(lg, *[x[2] for x in ghosts])
            rec_all = rec_all[np.argsort(gid_all, kind="stable")]
        srt, _ = oracle.sort(rec_all, g)
        start, end = oracle.findcells(srt, g)
        rs, rc = oracle.neighbourcells(srt, g, start, end)
        slot = srt[:, 3].view(np.int32)
        inv = np.empty_like(slot); inv[slot] = np.arange(slot.shape[0])       # slot -> sorted index
        ghost_sorted = inv[n:]
        for _it in range(K):
            lam, _ = oracle.calclambda(srt, rs, rc, P)
            f_lo, f_hi = exchange(rank, world, lam[inv[b_lo]], lam[inv[b_hi]])
            fresh = cat(f_lo, f_hi)
            if fresh is not None:
                lam[ghost_sorted] = fresh
            srt = oracle.updatepos(srt, rs, rc, lam, P, g)
            f_lo, f_hi = exchange(rank, world, srt[inv[b_lo]], srt[inv[b_hi]])
            fresh = cat(f_lo, f_hi)
            if fresh is not None:
                srt[ghost_sorted, :3] = fresh[:, :3]
        oracle.update(srt, P, pos_all, vel_all)
        lp, lv = pos_all[:n].copy(), vel_all[:n].copy()
    parts = [None] * world
    dist.all_gather_object(parts, (lp, lv, lg, migrated))
    ok = True
    if rank == 0:
        gp, gv = np.zeros_like(pos), np.zeros_like(vel)
        seen = np.zeros(pos.shape[0], int)
        for p_, v_, g_, _ in parts:
            gp[g_], gv[g_] = p_, v_
            seen[g_] += 1
        sim = oracle.Sim(pos.shape[0], g)
        for _ in range(steps):
            sim.step(pos, vel, P, K)
        dp, dv = np.max(np.abs(gp - pos)), np.max(np.abs(gv - vel))
        mig = sum(p[3] for p in parts)
        ok = bool(np.all(seen == 1) and dp < 1e-4 and dv < 1e-2 and mig > 0)
        if canonical:
            ok = ok and np.array_equal(gp.view(np.uint32), pos.view(np.uint32)) and np.array_equal(gv.view(np.uint32), vel.view(np.uint32))
        print("SLAB_MODEL ok=%s canonical=%s dp=%.3g dv=%.3g migrated=%d planes=%s" % (ok, canonical, dp, dv, mig, planes))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
