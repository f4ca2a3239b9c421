#!/usr/bin/env python
"""bench.py -- particle-steps/s of the per-timestep PBF path on synthetic dam-break scenes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one SPH::Run (predict, sort, cells, K_solver x (lambda, delta-p), update, vorticity+XSPH) over the whole
scene.  N=1 workload: BASELINE.json configs[2], the headline "dam-break 8M particles, 4 solver iters, full pipeline on
1xB200" (256x128x256 = 8,388,608 particles, grid 512x256x512, vorticity + XSPH on).  N>1: weak scaling, one such slab
per GPU (see DESIGN.md "Multi-GPU").  Prints ONE JSON line on rank 0.
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

METRIC = "particle-steps/s (4 solver iters, vorticity+XSPH) at 8M particles per GPU; HBM GB/s vs B200 peak"
UNIT = "particle-steps/s"
C3 = dict(n3=(256, 128, 256), grid=(512, 256, 512), iters=4, vort=1)
# algorithmic bytes per particle-step, SURVEY.md 8(d) / BASELINE.md section 3: 172 + 16 P + 56 K + 128 vort
STAGE_BYTES = {"lambda": 20, "delta_p": 36, "vorticity_a": 64, "vorticity_b": 64}


def algorithmic_bytes(grid, iters, vort):
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


def scene(n3, origin=(32.5, 0.5, 32.5)):
    import pbf_b200
    return pbf_b200.dam_break(*n3, origin=origin)


def cpu_oracle_rate(steps, warmup, iters, vort, grid):
    """The CPU oracle (oracle/pbf_oracle.c, all host threads) on a bounded sample of the C3 workload."""
    import oracle
    n3 = (256, 128, 32)           # 1,048,576 particles of the same lattice / grid / parameters
    pos, vel = oracle.dam_break(*n3)
    g = oracle.make_grid(*grid, ref_quirks=0)
    P = oracle.default_params()
    sim = oracle.Sim(pos.shape[0], g)
    for _ in range(warmup):
        sim.step(pos, vel, P, iters, vorticity=bool(vort))
    t0 = time.perf_counter()
    for _ in range(steps):
        sim.step(pos, vel, P, iters, vorticity=bool(vort))
    dt = time.perf_counter() - t0
    return pos.shape[0] * steps / dt, dt / steps, {
        "kind": "port", "cores": oracle.num_threads(),
        "sample": "dam-break %dx%dx%d = %d particles of the C3 lattice in the C3 grid, %d iters, vorticity %s, %d warm-up + %d timed steps"
                  % (n3 + (pos.shape[0], iters, "on" if vort else "off", warmup, steps))}


def run_reference(args):
    """--impl reference: the reference's GLSL cannot run here (no GL); the CPU oracle port stands in."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, sec, info = cpu_oracle_rate(args.steps, min(args.warmup, 1), C3["iters"], C3["vort"], C3["grid"])
    info["value"] = val
    info["unit"] = UNIT
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "dam-break 8M particles, 4 solver iters, full pipeline (bounded CPU sample, see cpu_baseline.sample)"},
        "cpu_baseline": info,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--small", action="store_true", help="debug: 1M particles instead of 8M")
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
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = dict(C3)
    if args.small:
        cfg = dict(n3=(128, 64, 128), grid=(256, 128, 256), iters=4, vort=1)
    if world > 1:
        from pbf_b200 import slab
        return slab.bench(args, cfg, rank, world, local, METRIC, UNIT, peaks, ClockSampler, algorithmic_bytes)

    pos, vel = scene(cfg["n3"])
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
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "dam-break %dx%dx%d = %d particles, grid %dx%dx%d, %d solver iters, vorticity+XSPH %s (BASELINE configs[2], headline)"
                               % (cfg["n3"] + (n,) + cfg["grid"] + (cfg["iters"], "on" if cfg["vort"] else "off")),
                   "l2": "working set (~1.6 GB of particle arrays + 512 MB cell table) far exceeds the 126 MB L2; no flush needed",
                   "ref_quirks": 0, "cuda_graph": True,
                   "step_algorithmic_bytes_per_particle": step_bytes,
                   "step_hbm_frac_of_peak": step_bytes * value / 1e9 / peak,
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
        val, sec, info = cpu_oracle_rate(3, 1, cfg["iters"], cfg["vort"], cfg["grid"])
        info.update({"value": val, "unit": UNIT})
        out["cpu_baseline"] = info
    print(json.dumps(out))


if __name__ == "__main__":
    main()
