#!/usr/bin/env python
"""bench.py -- particle-steps/s of the per-timestep PBF path on synthetic dam-break scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME]

One "step" = one SPH::Run (predict, sort, cells, K_solver x (lambda, delta-p), update, vorticity+XSPH) over the whole
scene.  Prints ONE JSON line on rank 0.  Workloads (BASELINE.json `configs`, sizes from SURVEY.md 8d):

  c1      configs[0]  32^3 = 32,768 particles, grid 128x64x128, 3 solver iters, vorticity off
  c2      configs[1]  128x64x128 = 1M particles, grid 256x128x256, 3 iters, vorticity + XSPH
  c3      configs[2]  256x128x256 = 8M particles, grid 512x256x512, 4 iters, vorticity + XSPH   <- default at N = 1 (headline)
  weak    weak scaling, one c3-sized block per GPU, BALLISTIC SPLASH: block r starts with v0 = (0, 10, +-20), sign
          alternating per slab, so particles cross the slab planes from the first step on       <- default at N > 1
  weak16  configs[4]  as `weak` with 256x128x512 = 16M particles per GPU                        (--per-gpu 16M)
  strong  configs[3]  512x256x512 = 64M-particle tank in a 1024x512x1024 grid cut into N z-slabs (--scaling strong)

`--scene rest|splash` overrides the initial velocities.  `--impl reference` times the CPU oracle port of the reference's
shaders (its GLSL cannot run here: no GL) with every host thread, on the same workload description.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

UNIT = "particle-steps/s"
METRIC = "particle-steps/s (4 solver iters, vorticity+XSPH) at 8M particles per GPU; HBM GB/s vs B200 peak"
SINGLE = {
    "c1": dict(n3=(32, 32, 32), grid=(128, 64, 128), iters=3, vort=0, label="BASELINE configs[0]"),
    "c2": dict(n3=(128, 64, 128), grid=(256, 128, 256), iters=3, vort=1, label="BASELINE configs[1]"),
    "c3": dict(n3=(256, 128, 256), grid=(512, 256, 512), iters=4, vort=1, label="BASELINE configs[2], headline"),
}
MULTI = {
    # per-GPU block and per-GPU grid depth (weak) / whole tank and whole grid (strong)
    "weak": dict(n3=(256, 128, 256), grid=(512, 256, 512), iters=4, vort=1, scene="splash", scaling="weak",
                 label="weak scaling, 8M particles per GPU, ballistic splash"),
    "weak16": dict(n3=(256, 128, 512), grid=(512, 256, 1024), iters=4, vort=1, scene="splash", scaling="weak",
                   label="BASELINE configs[4]: weak scaling, 16M particles per GPU, ballistic splash"),
    "strong": dict(n3=(512, 256, 512), grid=(1024, 512, 1024), iters=4, vort=1, scene="rest", scaling="strong",
                   label="BASELINE configs[3]: 64M-particle tank, strong scaling"),
}
SPLASH_V = (0.0, 10.0, 20.0)      # v0 = (0, 10, +-20), sign alternating per z-slab
# algorithmic bytes per particle-step, SURVEY.md 8(d) / BASELINE.md section 3: 172 + 16 P + 56 K + 128 vort
STAGE_BYTES = {"lambda": 20, "delta_p": 36, "vorticity_a": 64, "vorticity_b": 64}


def sort_passes(grid):
    import pbf_b200
    return pbf_b200.sort_passes(grid)


def algorithmic_bytes(grid, iters, vort):
    # P = ceil(keybits / 8) as SURVEY.md 8(d) defines it (the figure does not follow the implementation's digit width)
    import pbf_b200
    passes = (pbf_b200.sort_bits(grid) + 7) // 8
    return 172 + 16 * passes + 56 * iters + 128 * vort


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(kernel, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json, written by profiles/summarize.py); only valid for the particle count it was captured at."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p))
    e = t.get("kernels", {}).get(kernel)
    if not e or t.get("particles") != n:
        return None, None
    return e["dram_bytes_per_launch"], "ncu --set full, %s (bytes per launch)" % t.get("capture", "profiles/")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False
        self.q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        mhz = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i] == "Active" for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": int(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def workload_text(name, cfg, world, scene):
    n3, grid = cfg["n3"], cfg["grid"]
    if name in SINGLE:
        return ("dam-break %dx%dx%d = %d particles, grid %dx%dx%d, %d solver iters, vorticity+XSPH %s (%s)"
                % (n3 + (n3[0] * n3[1] * n3[2],) + grid + (cfg["iters"], "on" if cfg["vort"] else "off", cfg["label"])))
    if cfg["scaling"] == "weak":
        return ("%s: %dx%dx%d = %d particles per GPU x %d GPUs (one block %d x deeper along z), grid %dx%dx%d per GPU, "
                "%d solver iters, vorticity+XSPH %s, scene %s; z-slabs with 1-layer halos"
                % (cfg["label"], n3[0], n3[1], n3[2], n3[0] * n3[1] * n3[2], world, world, grid[0], grid[1], grid[2],
                   cfg["iters"], "on" if cfg["vort"] else "off", scene))
    return ("%s: %dx%dx%d = %d particles in a %dx%dx%d grid on %d GPUs, %d solver iters, vorticity+XSPH %s, scene %s; "
            "z-slabs with 1-layer halos" % (cfg["label"], n3[0], n3[1], n3[2], n3[0] * n3[1] * n3[2], grid[0], grid[1],
                                            grid[2], world, cfg["iters"], "on" if cfg["vort"] else "off", scene))


def static_config(name, cfg, world, scene):
    """What both arms print as `config`: a description of the workload, nothing measured."""
    return {"workload": workload_text(name, cfg, world, scene), "name": name, "scene": scene,
            "parallelism": "single" if world == 1 else "slab%d" % world,
            "l2": "working set (>= 100 B per particle + 16 B per cell: GBs) far exceeds the 126 MB L2; no flush needed",
            "ref_quirks": 0}


def pick_config(args, world):
    name = args.config
    if args.scaling == "strong":
        name = "strong"
    if args.per_gpu:
        if args.per_gpu.upper() != "16M":
            raise SystemExit("bench.py: --per-gpu supports 16M (BASELINE configs[4])")
        name = "weak16"
    if name is None:
        name = "c3" if world == 1 else "weak"
    if args.small:                                                    # debug sizes
        name = {"c3": "c2", "weak": "weak"}.get(name, name)
    cfg = dict(SINGLE[name]) if name in SINGLE else dict(MULTI[name])
    if args.small and name in MULTI:
        cfg["n3"] = tuple(max(32, v // 4) for v in cfg["n3"])
        cfg["grid"] = tuple(max(64, v // 4) for v in cfg["grid"])
    scene = args.scene or cfg.get("scene", "rest")
    return name, cfg, scene


def rank_block(name, cfg, rank, world, scene):
    """This rank's particles: (pos, vel, gid, z_planes, grid_global).  Single configs: the whole block."""
    from pbf_b200 import slab
    n3, grid = cfg["n3"], cfg["grid"]
    if name in SINGLE or (world == 1 and cfg.get("scaling") == "strong"):
        import pbf_b200
        pos, vel = pbf_b200.dam_break(*n3)
        gid = np.arange(pos.shape[0], dtype=np.uint32)
        planes, ggrid = [0, grid[2]], grid
    elif cfg["scaling"] == "weak":
        pos, vel, gid, planes, gz = slab.weak_scene(rank, world, n3, grid[2])
        ggrid = (grid[0], grid[1], gz)
    else:                                                             # strong: the tank's lattice layers split evenly
        assert n3[2] % world == 0, "tank depth must divide by the GPU count"
        pos, vel, gid, planes, _ = slab.weak_scene(rank, world, (n3[0], n3[1], n3[2] // world), grid[2] // world)
        planes[-1] = grid[2]
        ggrid = grid
    if scene == "splash":
        vel[:, 0], vel[:, 1] = SPLASH_V[0], SPLASH_V[1]
        vel[:, 2] = SPLASH_V[2] if rank % 2 == 0 else -SPLASH_V[2]
    return pos, vel, gid, planes, ggrid


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle_rate(name, cfg, scene, steps, warmup, budget_s):
    """The CPU oracle (oracle/pbf_oracle.c) with every host thread on (a bounded sample of) the workload: the whole
    single-GPU block when `warmup + steps` of it fit the time budget, else its first 1/8 of the lattice layers."""
    import oracle
    oracle.set_num_threads(host_threads())     # torchrun exports OMP_NUM_THREADS=1; the baseline uses all cores regardless
    n3, grid = cfg["n3"], cfg["grid"]
    if name in MULTI and cfg["scaling"] == "strong":
        n3, grid = (n3[0], n3[1], n3[2] // 8), (grid[0], grid[1], grid[2] // 8 + 64)
        what = "one 8-GPU slab's share (1/8 of the tank's layers)"
    elif name in MULTI:
        what = "one GPU's block"
    else:
        what = "the whole scene"
    P = oracle.default_params()

    def make(n3s):
        pos, vel = oracle.dam_break(*n3s)
        if scene == "splash":
            vel[:, :3] = SPLASH_V
        return pos, vel, oracle.Sim(pos.shape[0], oracle.make_grid(*grid, ref_quirks=0))

    probe3 = (n3[0], n3[1], max(8, n3[2] // 8))
    pos, vel, sim = make(probe3)
    t0 = time.perf_counter()
    sim.step(pos, vel, P, cfg["iters"], vorticity=bool(cfg["vort"]))
    per_particle = (time.perf_counter() - t0) / pos.shape[0]
    full = per_particle * n3[0] * n3[1] * n3[2] * (steps + warmup) <= budget_s
    used3 = n3 if full else probe3
    if full and used3 != probe3:
        del sim
        pos, vel, sim = make(used3)
        done = 0
    else:
        done = 1                                # the probe step was the first warm-up step
    for _ in range(max(0, warmup - done)):
        sim.step(pos, vel, P, cfg["iters"], vorticity=bool(cfg["vort"]))
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step(pos, vel, P, cfg["iters"], vorticity=bool(cfg["vort"]))
    dt = time.perf_counter() - t0
    n = pos.shape[0]
    return n * steps / dt, dt / steps, {
        "kind": "port", "cores": oracle.num_threads(),
        "sample": "%s%s: dam-break %dx%dx%d = %d particles in a %dx%dx%d grid, %d iters, vorticity %s, scene %s, %d warm-up + %d timed steps"
                  % (what, "" if full else " cut to its first 1/8 of the lattice layers (time budget)", used3[0], used3[1],
                     used3[2], n, grid[0], grid[1], grid[2], cfg["iters"], "on" if cfg["vort"] else "off", scene, warmup, steps)}


def run_reference(args):
    """--impl reference: the reference's GLSL cannot run here (no GL); the CPU oracle port stands in."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    name, cfg, scene = pick_config(args, world)
    val, sec, info = cpu_oracle_rate(name, cfg, scene, args.steps, args.warmup, budget_s=240.0)
    info["value"] = val
    info["unit"] = UNIT
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": cfg.get("scaling", "weak"),
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": static_config(name, cfg, world, scene),
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_single(args, name, cfg, scene, local):
    import torch
    import pbf_b200
    pos, vel, _, _, _ = rank_block(name, cfg, 0, 1, scene)
    n = pos.shape[0]
    sph = pbf_b200.SPH(n, cfg["grid"], ref_quirks=False, device=local)
    sph.SetNumSolverIterations(cfg["iters"])
    sph.SetVorticityConfinementEnabled(bool(cfg["vort"]))
    sph.upload(pos, vel)
    stream = torch.cuda.ExternalStream(sph.stream, device=local)

    def timed(fn, reps):
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

    # ---- whole step, state resident in HBM ---------------------------------------------------------------------------
    sph.Run(args.warmup)
    sph.sync()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = sph.kernel_launches
    ms_step = timed(lambda: sph.Run(1), args.steps)
    launches = sph.kernel_launches - l0
    sampler.stop_flag = True
    sampler.join()
    value = n / (ms_step * 1e-3)

    # ---- per-kernel durations ------------------------------------------------------------------------------------------
    # (i) the two density-constraint kernels over the SAME steps as the timed region (the scene is restarted and warmed
    #     up again; a sweep gets slower as the lattice disorders, so the step range matters): timing mode records a CUDA
    #     event before every solver kernel on the library's stream (pbf_get_solver_kernel_timings);
    # (ii) the stage entry points alone, for the kernels the library has no in-step events for.
    sph.upload(pos, vel)
    sph.Run(args.warmup)
    sph.enable_timing(True)
    in_step = {"lambda": [], "delta_p": []}
    phase_acc = []
    for _ in range(args.steps):
        sph.Run(1)
        a, b = sph.get_solver_kernel_timings()
        in_step["lambda"].append(a); in_step["delta_p"].append(b)
        phase_acc.append(sph.get_timings())
    sph.enable_timing(False)
    phases = [float(x) for x in np.mean(np.array(phase_acc), axis=0)]
    kernel_ms = {k: float(np.mean(v)) for k, v in in_step.items()}
    sph.predict(); sph.sort(); sph.build_cells()
    tiles, tiled = sph.tile_stats()
    stage_ms = {}
    reps = 5
    stage_ms["lambda"] = timed(sph.calc_lambda, reps)
    stage_ms["delta_p"] = timed(sph.update_positions, reps)
    sph.calc_lambda(); sph.finalize()
    if cfg["vort"]:
        stage_ms["vorticity_a+b"] = timed(sph.vorticity, reps)
    sph.upload(pos, vel)
    sph.Run(3)
    peak, peak_src = peaks()
    dom = max(("lambda", "delta_p"), key=lambda k: kernel_ms[k])
    dom_bytes = STAGE_BYTES[dom] * n
    achieved = dom_bytes / (kernel_ms[dom] * 1e-3) / 1e9
    step_bytes = algorithmic_bytes(cfg["grid"], cfg["iters"], cfg["vort"])
    traffic, traffic_src = ncu_traffic("k_" + dom, n)

    # ---- end to end through the public call with HOST buffers -------------------------------------------------------------
    hp = torch.from_numpy(pos).pin_memory()
    hv = torch.from_numpy(vel).pin_memory()
    sph.step_host(hp, hv, 1)
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        sph.step_host(hp, hv, 1)
    e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3

    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": cfg.get("scaling", "weak"), "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": static_config(name, cfg, 1, scene),
        "detail": {"cuda_graph": True, "sort_passes": sort_passes(cfg["grid"]),
                   "step_algorithmic_bytes_per_particle": step_bytes,
                   "step_hbm_frac_of_peak": step_bytes * value / 1e9 / peak,
                   "other_kernels_hbm_frac": {k: STAGE_BYTES.get(k, 0) * n / (v * 1e-3) / 1e9 / peak for k, v in kernel_ms.items()},
                   "phase_ms": dict(zip(["predict", "sort", "neighbour_cells", "solver", "vorticity"], phases)),
                   "kernel_ms_in_step": kernel_ms, "ms_per_step_timing_mode": float(sum(phases)),
                   "stage_ms_alone": stage_ms, "tiles": tiles, "tiles_on_tiled_path": tiled},
        "clocks": sampler.summary(),
        "gpu_launches": int(launches),
        "e2e": {"value": n / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 2 * n * 16, "d2h_bytes_per_step": 2 * n * 16,
                "ms_per_step": e2e_ms, "call": "pbf_step_host (pinned host pos+vel in, pos+vel out)"},
        "roofline": {"bound": "hbm", "kernel": "k_" + dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": kernel_ms[dom],
                     "launch_ms_source": "mean over every launch of the kernel in the same steps as the timed region (scene restarted), CUDA events on the library stream around each solver kernel",
                     "note": "density-constraint kernels are FP32-issue bound, not HBM bound (DESIGN.md); frac is reported against HBM as the north star asks"},
    }
    if not args.no_cpu_baseline:
        val, sec, info = cpu_oracle_rate(name, cfg, scene, 3, 1, budget_s=30.0)
        info.update({"value": val, "unit": UNIT})
        out["cpu_baseline"] = info
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default=None, choices=sorted(SINGLE) + sorted(MULTI))
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="strong = BASELINE configs[3] (64M tank)")
    ap.add_argument("--per-gpu", default=None, help="16M = BASELINE configs[4]")
    ap.add_argument("--scene", default=None, choices=["rest", "splash"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small", action="store_true", help="debug: quarter-size scenes")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import pbf_b200
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; pbf_b200 has no CPU path")
    torch.cuda.set_device(local)
    name, cfg, scene = pick_config(args, world)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        from pbf_b200 import slab
        return slab.bench(args, name, cfg, scene, rank, world, local, sys.modules[__name__])
    run_single(args, name, cfg, scene, local)


if __name__ == "__main__":
    main()
