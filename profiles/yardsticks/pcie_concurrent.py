#!/usr/bin/env python
"""Yardstick for the end-to-end lines: what the host <-> device copies of pbf_slab_step_host cost by
themselves when k GPUs of the box copy at the same time -- 268 MB (8M particles x 2 float4 arrays) from pinned host memory
to the device and back, per GPU, k = 1, 2, 4, 8.  Not part of the product; prints one JSON line.

  python profiles/yardsticks/pcie_concurrent.py > gpurun_out/pcie_concurrent.json
"""
import json
import subprocess
import time

import torch

MB = 268435456
n = torch.cuda.device_count()
host = [torch.empty(MB, dtype=torch.uint8).pin_memory() for _ in range(n)]
dev = [torch.empty(MB, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]
for h in host:
    h.fill_(1)
out = {"bytes_per_gpu_per_direction": MB, "gpus": n, "runs": []}
for k in [1, 2, 4, 8]:
    if k > n:
        break
    res = {}
    for what in ("h2d", "d2h", "h2d_then_d2h"):
        best = 1e9
        for rep in range(4):
            for i in range(k):
                torch.cuda.synchronize(i)
            t0 = time.perf_counter()
            for i in range(k):
                with torch.cuda.stream(streams[i]):
                    if what != "d2h":
                        dev[i].copy_(host[i], non_blocking=True)
                    if what != "h2d":
                        host[i].copy_(dev[i], non_blocking=True)
            for i in range(k):
                streams[i].synchronize()
            best = min(best, time.perf_counter() - t0)
        nb = MB * k * (2 if what == "h2d_then_d2h" else 1)
        res[what] = {"ms": round(best * 1e3, 3), "aggregate_GBps": round(nb / best / 1e9, 1), "per_gpu_GBps": round(nb / best / 1e9 / k, 1)}
    out["runs"].append({"concurrent_gpus": k, **res})
try:
    out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=30).stdout.splitlines()[:14]
except Exception as e:  # noqa: BLE001
    out["topo"] = str(e)
print(json.dumps(out))
