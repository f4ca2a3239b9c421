"""Host runtime of the slab decomposition (multi GPU): partitioning, NCCL bootstrap over torch.distributed, and the
weak-scaling benchmark.  The device side lives in csrc/slab.cu behind the pbf_slab_* entry points of include/pbf_c.h.

One process per GPU (torchrun).  torch.distributed is plumbing only: it broadcasts the 128-byte NCCL unique id and
reduces timings; the halo/migration traffic itself goes through the library's own communicator on its own stream.
"""
import ctypes as C
import os
import json
import time

import numpy as np

from . import SPH, _check, _ptr, dam_break, lib


def plan_slabs(cell_z, gz_global, nranks):
    """Cut [0, gz_global) into nranks contiguous ranges of whole cell layers holding about the same number of particles.

    cell_z: integer cell layer (int(clamp(z))) of every particle.  Returns z_planes (nranks + 1 ints); every slab
    owns at least 2 layers (the halo scheme needs boundary layers z_lo and z_hi-1 to be distinct)."""
    cell_z = np.asarray(cell_z)
    hist = np.bincount(np.clip(cell_z, 0, gz_global - 1), minlength=gz_global)
    cum = np.concatenate([[0], np.cumsum(hist)])
    planes = [0]
    for r in range(1, nranks):
        target = cum[-1] * r / nranks
        z = int(np.searchsorted(cum, target, side="left"))
        z = max(z, planes[-1] + 2)
        z = min(z, gz_global - 2 * (nranks - r))
        planes.append(z)
    planes.append(gz_global)
    if any(b - a < 2 for a, b in zip(planes[:-1], planes[1:])):
        raise ValueError("domain too shallow for %d slabs of >= 2 layers" % nranks)
    return planes


def rebalance_planes(layer_counts, old_planes, max_shift=1, max_thickness=None):
    """New slab planes from the particle count per global cell layer: the equal-share planes of plan_slabs, but every inner
    plane moves at most `max_shift` layers per call (the particles that change owner travel through ONE step's migration),
    slabs keep at least 2 layers and at most `max_thickness` (what the handles' grids were allocated for)."""
    counts = np.asarray(layer_counts, dtype=np.int64)
    n = len(old_planes) - 1
    limit = None if max_thickness is None else ([int(max_thickness)] * n if np.isscalar(max_thickness) else [int(t) for t in max_thickness])
    cum = np.concatenate([[0], np.cumsum(counts)])
    new = [int(old_planes[0])]
    for r in range(1, n):
        want = int(np.searchsorted(cum, cum[-1] * r / n, side="left"))
        z = int(np.clip(want, old_planes[r] - max_shift, old_planes[r] + max_shift))
        z = max(z, new[-1] + 2)
        z = min(z, int(old_planes[-1]) - 2 * (n - r))
        if limit is not None:
            z = min(z, new[-1] + limit[r - 1])                      # slab r-1 may not outgrow its window ...
        new.append(z)
    new.append(int(old_planes[-1]))
    if limit is not None and any(b - a > t for a, b, t in zip(new[:-1], new[1:], limit)):
        return [int(z) for z in old_planes]            # ... nor slab r: if the windows cannot hold the new cut, keep the planes
    return new


def cell_layer(pos, gz_global):
    return np.clip(pos[:, 2], 0.0, float(gz_global)).astype(np.int32)


def split_scene(pos, vel, z_planes, gz_global):
    """Global scene -> per-rank (pos, vel, gid); gid = index in the global arrays."""
    cz = cell_layer(pos, gz_global)
    out = []
    for r in range(len(z_planes) - 1):
        m = np.nonzero((cz >= z_planes[r]) & (cz < z_planes[r + 1]))[0].astype(np.uint32)
        out.append((np.ascontiguousarray(pos[m]), np.ascontiguousarray(vel[m]), m))
    return out


class SlabSPH(SPH):
    """One rank's slab: an SPH handle whose cell tables cover layers [z_lo-1, z_hi+1) of the global domain."""

    def __init__(self, rank, nranks, z_planes, grid_xy, gz_global, capacity, halo_capacity, wall=(16.0, 0.0, 16.0),
                 device=-1, extra_layers=0):
        """extra_layers: how many layers the window may grow beyond its initial thickness (set_planes / rebalancing)."""
        self.rank, self.nranks = rank, nranks
        self.z_lo, self.z_hi = int(z_planes[rank]), int(z_planes[rank + 1])
        self.gz_global = gz_global
        self.halo_capacity = halo_capacity
        self.max_thickness = self.z_hi - self.z_lo + int(extra_layers)
        capacity = (capacity + 511) // 512 * 512
        super().__init__(capacity, (grid_xy[0], grid_xy[1], self.max_thickness + 2), wall=wall, ref_quirks=False,
                         device=device, use_graph=False, capacity=capacity)
        self.capacity = capacity

    def init_nccl(self, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        _check(lib().pbf_slab_init(self._h, buf, self.rank, self.nranks, self.z_lo, self.z_hi, self.gz_global,
                                   self.halo_capacity))

    def connect_p2p(self, dist, device):
        """Peer-memory halo refresh: all-gather the ranks' 64-byte CUDA IPC mailbox handles over torch.distributed and
        open the two neighbours'.  PBF_SLAB_P2P=0 keeps the NCCL send/recv path (the baseline it is measured against)."""
        import torch
        if os.environ.get("PBF_SLAB_P2P", "1") == "0" or self.nranks == 1:
            return False
        mine = (C.c_char * 64)()
        _check(lib().pbf_slab_p2p_handle(self._h, mine))
        t = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8).clone().to(device)
        parts = [torch.empty_like(t) for _ in range(self.nranks)]
        dist.all_gather(parts, t)
        raw = [bytes(x.cpu().numpy().tobytes()) for x in parts]
        lo = (C.c_char * 64).from_buffer_copy(raw[self.rank - 1]) if self.rank > 0 else None
        hi = (C.c_char * 64).from_buffer_copy(raw[self.rank + 1]) if self.rank + 1 < self.nranks else None
        _check(lib().pbf_slab_p2p_connect(self._h, lo, hi))
        dist.barrier()        # nobody steps before every mailbox is mapped
        return True

    def upload_slab(self, pos, vel, gid):
        pos = np.ascontiguousarray(pos, np.float32)
        vel = np.ascontiguousarray(vel, np.float32)
        gid = np.ascontiguousarray(gid, np.uint32)
        _check(lib().pbf_slab_upload(self._h, _ptr(pos), _ptr(vel), _ptr(gid), pos.shape[0]))

    def download_slab(self):
        n = C.c_uint32()
        _check(lib().pbf_slab_download(self._h, None, None, None, C.byref(n)))
        pos = np.empty((n.value, 4), np.float32)
        vel = np.empty((n.value, 4), np.float32)
        gid = np.empty(n.value, np.uint32)
        _check(lib().pbf_slab_download(self._h, _ptr(pos), _ptr(vel), _ptr(gid), C.byref(n)))
        return pos, vel, gid

    def download_highlight(self):
        n = C.c_uint32()
        _check(lib().pbf_slab_download(self._h, None, None, None, C.byref(n)))
        hl = np.empty(n.value, np.uint32)
        _check(lib().pbf_slab_download_highlight(self._h, _ptr(hl)))
        return hl

    def Run(self, nsteps=1):
        _check(lib().pbf_slab_step(self._h, nsteps))

    def step_host(self, pos, vel, gid, n, capacity, nsteps=1):
        """End-to-end call: n local particles from (pinned) host arrays in, nsteps, the particles this rank owns afterwards
        back out into the same arrays; returns their number."""
        m = C.c_uint32()
        _check(lib().pbf_slab_step_host(self._h, _ptr(pos), _ptr(vel), _ptr(gid), n, capacity, C.byref(m), nsteps))
        return m.value

    def layer_counts(self):
        """This rank's particles per GLOBAL cell layer (what rebalancing sums over the ranks)."""
        out = np.zeros(self.gz_global, np.uint32)
        _check(lib().pbf_slab_layer_counts(self._h, _ptr(out)))
        return out

    def set_planes(self, z_lo, z_hi):
        _check(lib().pbf_slab_set_planes(self._h, int(z_lo), int(z_hi)))
        self.z_lo, self.z_hi = int(z_lo), int(z_hi)

    def rebalance(self, dist, z_planes, max_shift=1):
        """Real ranks: sum the layer histograms over torch.distributed, move every inner plane by at most max_shift layers
        towards equal particle counts, adopt this rank's new planes.  Collective: every rank calls it between steps.
        Returns the new planes (the same list on every rank)."""
        import torch
        t = torch.from_numpy(self.layer_counts().astype(np.int64))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
        thick = torch.zeros(self.nranks, dtype=torch.int64, device=t.device)
        thick[self.rank] = self.max_thickness
        dist.all_reduce(thick)
        new = rebalance_planes(t.cpu().numpy(), z_planes, max_shift, thick.cpu().tolist())
        self.set_planes(new[self.rank], new[self.rank + 1])
        return new

    def phase_times(self):
        ms = (C.c_float * 5)()
        _check(lib().pbf_slab_phase_times(self._h, ms))
        return dict(zip(("predict_migrate", "arrivals_ghosts_out", "ghosts_in_sort_cells", "solver", "vorticity"), [float(x) for x in ms]))

    def stats(self):
        out = (C.c_uint64 * 8)()
        _check(lib().pbf_slab_stats(self._h, out))
        keys = ("n_local", "ghosts_lo", "ghosts_hi", "boundary_lo", "boundary_hi", "migrated", "exchanges", "bytes_sent")
        return dict(zip(keys, [int(v) for v in out]))


def unique_id():
    buf = (C.c_char * 128)()
    _check(lib().pbf_slab_unique_id(buf))
    return bytes(buf)


def broadcast_unique_id(dist, rank, device=None):
    """Rank 0 creates the NCCL unique id, everybody receives it through torch.distributed (any backend)."""
    import torch
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(unique_id()), dtype=torch.uint8).clone()
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


class VirtualGroup:
    """nranks slabs inside ONE process on one GPU, exchanged with device copies (tests of the slab logic)."""

    def __init__(self, pos, vel, nranks, grid, wall=(16.0, 0.0, 16.0), halo_capacity=1 << 16, slack=1.5, device=-1,
                 extra_layers=0, z_planes=None):
        gz_global = grid[2]
        self.z_planes = list(z_planes) if z_planes is not None else plan_slabs(cell_layer(pos, gz_global), gz_global, nranks)
        parts = split_scene(pos, vel, self.z_planes, gz_global)
        self.ranks = []
        for r, (p, v, g) in enumerate(parts):
            cap = int(max(p.shape[0], 512) * slack) + 2 * halo_capacity
            self.ranks.append(SlabSPH(r, nranks, self.z_planes, grid[:2], gz_global, cap, halo_capacity, wall, device,
                                      extra_layers=extra_layers))
        hs = (C.c_void_p * nranks)(*[s._h for s in self.ranks])
        _check(lib().pbf_slab_init_group(hs, nranks, (C.c_int32 * (nranks + 1))(*self.z_planes), gz_global, halo_capacity))
        for s, (p, v, g) in zip(self.ranks, parts):
            s.upload_slab(p, v, g)
        self.n = pos.shape[0]

    def set_params(self, **kw):
        for s in self.ranks:
            s._set(**kw)

    def set_canonical_order(self, on=True):
        for s in self.ranks:
            s.set_canonical_order(on)

    def Run(self, nsteps=1):
        self.ranks[0].Run(nsteps)

    def gather(self):
        pos = np.zeros((self.n, 4), np.float32)
        vel = np.zeros((self.n, 4), np.float32)
        seen = np.zeros(self.n, np.int32)
        for s in self.ranks:
            p, v, g = s.download_slab()
            pos[g], vel[g] = p, v
            seen[g] += 1
        assert np.all(seen == 1), "every particle must be owned by exactly one rank"
        return pos, vel

    def rebalance(self, max_shift=1):
        """Move the inner planes by at most max_shift layers towards equal particle counts (between steps)."""
        counts = sum(s.layer_counts().astype(np.int64) for s in self.ranks)
        new = rebalance_planes(counts, self.z_planes, max_shift, [s.max_thickness for s in self.ranks])
        for r, s in enumerate(self.ranks):
            s.set_planes(new[r], new[r + 1])
        self.z_planes = new
        return new

    def toggle_highlight(self, gids):
        """Simulation::OnMouseDown's highlight toggle for particles given by GLOBAL id (the owner's slot is looked up)."""
        want = set(int(g) for g in np.atleast_1d(gids))
        for s in self.ranks:
            _, _, gid = s.download_slab()
            for slot in np.nonzero(np.isin(gid, list(want)))[0]:
                s.toggle_highlight(int(slot))

    def gather_highlight(self):
        out = np.zeros(self.n, np.uint32)
        for s in self.ranks:
            _, _, gid = s.download_slab()
            out[gid] = s.download_highlight()
        return out

    def close(self):
        for s in reversed(self.ranks):
            s.close()


def weak_scene(rank, nranks, n3, grid_local_z=512, origin=(32.5, 0.5, 32.5), spacing=0.94):
    """Rank's share of a dam-break block of n3[0] x n3[1] x (n3[2]*nranks) particles in a domain nranks*grid_local_z deep.

    Rank r generates lattice layers k in [n3[2]*r, n3[2]*(r+1)) -- exactly n3[0]*n3[1]*n3[2] particles, no overlap --
    and the slab planes are the cell layers those ranges start in.  The few particles of the last lattice layer that
    fall on the far side of a plane are handed over by the runtime's migration in the first step."""
    gz_global = grid_local_z * nranks
    planes = [0] + [int(origin[2] + spacing * n3[2] * r) for r in range(1, nranks)] + [gz_global]
    k0 = n3[2] * rank
    pos, vel = dam_break(n3[0], n3[1], n3[2], origin=(origin[0], origin[1], origin[2] + spacing * k0), spacing=spacing,
                         seed=12345 + rank, id0=0)
    j = np.arange(pos.shape[0], dtype=np.int64)          # dam_break walks x, z, y with y innermost
    x, zk, y = j // (n3[2] * n3[1]), (j // n3[1]) % n3[2], j % n3[1]
    gid = ((x * (n3[2] * nranks) + (k0 + zk)) * n3[1] + y).astype(np.uint32)
    return pos, vel, gid, planes, gz_global


def bench(args, name, cfg, scene, rank, world, local, B):
    """Multi-GPU lines of bench.py (B = the bench module): weak scaling (one block per GPU, default: ballistic splash) or
    strong scaling (one tank cut into `world` slabs).  Device time = CUDA events on the library's stream bracketed by
    barriers, max over ranks; end to end = pbf_slab_step_host with persistent pinned host buffers."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    pos, vel, gid, planes, ggrid = B.rank_block(name, cfg, rank, world, scene)
    n0 = pos.shape[0]
    halo_cap = 1 << 19 if n0 > (12 << 20) else 1 << 18
    s = SlabSPH(rank, world, planes, ggrid[:2], ggrid[2], int(n0 * 1.25) + 2 * halo_cap, halo_cap, device=local)
    s.init_nccl(broadcast_unique_id(dist, rank, dev))
    p2p = s.connect_p2p(dist, dev)
    s.SetNumSolverIterations(cfg["iters"])
    s.SetVorticityConfinementEnabled(bool(cfg["vort"]))
    s.upload_slab(pos, vel, gid)
    stream = torch.cuda.ExternalStream(s.stream, device=local)
    s.Run(args.warmup)
    s.sync()
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = s.kernel_launches
    m0 = s.stats()["migrated"]
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        s.Run(args.steps)
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ms = e0.elapsed_time(e1) / args.steps
    launches = s.kernel_launches - l0
    st = s.stats()
    phases = s.phase_times() if os.environ.get("PBF_SLAB_PHASES") == "1" else None
    t = torch.tensor([ms, float(st["n_local"]), float(st["migrated"] - m0), float(st["ghosts_lo"] + st["ghosts_hi"])],
                     dtype=torch.float64, device=dev)
    tmax = t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = t.clone()
    dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    tsum = t.clone()
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    # end to end: the rank's particles in from pinned host arrays, one step, the particles it then owns back out
    cap = s.capacity
    hp = torch.zeros((cap, 4), dtype=torch.float32).pin_memory()
    hv = torch.zeros((cap, 4), dtype=torch.float32).pin_memory()
    hg = torch.zeros((cap,), dtype=torch.int32).pin_memory()
    lp, lv, lg = s.download_slab()
    n_loc = lp.shape[0]
    hp[:n_loc] = torch.from_numpy(lp); hv[:n_loc] = torch.from_numpy(lv); hg[:n_loc] = torch.from_numpy(lg.view(np.int32))
    del lp, lv, lg
    n_loc = s.step_host(hp, hv, hg, n_loc, cap)                   # warm
    e2e_steps = max(3, min(args.steps, 10))
    moved = 0
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        moved += 2 * n_loc * 36
        n_loc = s.step_host(hp, hv, hg, n_loc, cap)
    dist.barrier()
    e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps * 1e3, moved / (2.0 * e2e_steps)], dtype=torch.float64, device=dev)
    e2e_max = e2e.clone()
    dist.all_reduce(e2e_max, op=dist.ReduceOp.MAX)
    dist.all_reduce(e2e, op=dist.ReduceOp.SUM)
    if rank == 0:
        sampler.stop_flag = True
        sampler.join()
        n_total = tsum[1].item()
        ms_step = tmax[0].item()
        value = n_total / (ms_step * 1e-3)
        peak, peak_src = B.peaks()
        local_grid = (ggrid[0], ggrid[1], max(b - a for a, b in zip(planes[:-1], planes[1:])) + 2)
        step_bytes = B.algorithmic_bytes(local_grid, cfg["iters"], cfg["vort"])
        per_gpu_gbs = step_bytes * value / world / 1e9
        print(json.dumps({
            "metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": B.static_config(name, cfg, world, scene),
            "detail": {"halo_transport": (("peer-memory stores + release flags over NVLink for everything (migration and ghost records into the neighbour's inbox; lambda, positions, |omega| into its mailbox); counts stay on the device, the step is one CUDA graph"
                                           if os.environ.get("PBF_SLAB_DEVCOUNT", "1") != "0" and os.environ.get("PBF_SLAB_FUSED") != "1" else
                                           "peer-memory stores + flags for lambda, positions, |omega|; NCCL send/recv for migration and ghost records, counts read back twice per step")
                                          if p2p else "NCCL send/recv"),
                       "particles_total": int(n_total), "particles_per_rank_min_max": [int(tmin[1].item()), int(tmax[1].item())],
                       "migrated_particles_in_timed_steps": int(tsum[2].item()),
                       "migrated_fraction_per_step": tsum[2].item() / max(1.0, n_total * args.steps),
                       "ghost_particles_total": int(tsum[3].item()),
                       "ms_per_step_min_over_ranks": tmin[0].item(),
                       "exchanges_per_step": 2 + 2 * cfg["iters"] + (1 if cfg["vort"] else 0),
                       "step_algorithmic_bytes_per_particle": step_bytes,
                       "step_hbm_frac_of_peak": per_gpu_gbs / peak, "rank0_phase_ms_last_step": phases},
            "clocks": sampler.summary(), "gpu_launches": int(launches),
            "e2e": {"value": n_total / (e2e_max[0].item() * 1e-3), "unit": B.UNIT, "h2d_bytes_per_step": int(e2e[1].item()),
                    "d2h_bytes_per_step": int(e2e[1].item()), "ms_per_step": e2e_max[0].item(),
                    "call": "pbf_slab_step_host per rank (persistent pinned host pos+vel+gid in, pos+vel+gid out)"},
            "roofline": {"bound": "hbm", "kernel": "whole step (per GPU)", "achieved": per_gpu_gbs,
                         "peak": peak, "unit": "GB/s", "frac": per_gpu_gbs / peak, "traffic": None,
                         "peak_source": peak_src, "note": "per-GPU algorithmic bytes / step time; see the N=1 line for the dominant kernel"},
        }))
    dist.barrier()
    s.close()
    dist.destroy_process_group()
