"""Real multi-GPU parity: needs >= 2 GPUs, otherwise skipped.  Both halo transports -- peer-memory stores with flags over
NVLink (the default) and NCCL send/recv (PBF_SLAB_P2P=0) -- against a single-domain run.  The single-GPU "virtual rank"
test in test_gpu_parity.py covers the same slab logic inside one process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# peer memory with device-side counts and the refreshes inside the sweeps (default); the same with a push and a pull kernel per
# refresh; the default again in canonical order (pbf_set_canonical_order on both sides: the slab run must equal the single-domain
# run BIT FOR BIT); the count read-back step with the push fused into the sweeps; NCCL
@pytest.mark.parametrize("p2p", ["1", "1k", "1c", "1f", "0"])
@pytest.mark.parametrize("world", [2, 4])
def test_slabs_match_single_domain(built_lib, world, p2p):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29617 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PBF_SLAB_P2P=p2p[0], PBF_SLAB_FUSED="1" if p2p.endswith("f") else "0",
                                PBF_SLAB_OVERLAP="0" if p2p.endswith("k") else "1",
                                PBF_TEST_CANONICAL="1" if p2p.endswith("c") else "0"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_RESULT ok=True canonical=%s p2p=%s" % (p2p.endswith("c"), p2p[0] == "1") in r.stdout, r.stdout[-3000:]
    print([l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT")][0])      # kept in the committed pytest -s logs
