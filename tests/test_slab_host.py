"""Host-side logic of the multi-GPU path on CPU: slab planning/splitting, and a world_size-2 gloo run of the CPU model
of the slab algorithm (tests/slab_model_worker.py) against the single-domain oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_and_split(built_lib):
    import oracle
    from pbf_b200 import slab
    pos, vel = oracle.dam_break(16, 8, 64, origin=(18.5, 0.5, 18.5))
    for nranks in (1, 2, 3, 4, 8):
        planes = slab.plan_slabs(slab.cell_layer(pos, 128), 128, nranks)
        assert planes[0] == 0 and planes[-1] == 128 and len(planes) == nranks + 1
        assert all(b - a >= 2 for a, b in zip(planes[:-1], planes[1:]))
        parts = slab.split_scene(pos, vel, planes, 128)
        counts = [p[0].shape[0] for p in parts]
        assert sum(counts) == pos.shape[0]
        assert max(counts) - min(counts) <= 2 * 16 * 8 * 2          # balanced to within a couple of lattice layers
        gids = np.concatenate([p[2] for p in parts])
        assert np.array_equal(np.sort(gids), np.arange(pos.shape[0]))
    with pytest.raises(ValueError):
        slab.plan_slabs(np.zeros(10, int), 6, 4)


def test_weak_scene_partitions_are_disjoint(built_lib):
    from pbf_b200 import slab
    n3 = (8, 4, 16)
    total, seen = 0, set()
    for r in range(3):
        p, v, gid, planes, gzg = slab.weak_scene(r, 3, n3, 64)
        cz = slab.cell_layer(p, gzg)
        assert np.all((cz >= planes[r] - 1) & (cz < planes[r + 1] + 1))     # stragglers migrate in step 1
        assert p.shape[0] == n3[0] * n3[1] * n3[2]
        assert not (seen & set(gid.tolist()))
        seen |= set(gid.tolist())
        total += p.shape[0]
    assert total == n3[0] * n3[1] * n3[2] * 3


@pytest.mark.parametrize("canonical", [False, True])
def test_slab_model_world2_gloo(built_lib, canonical):
    """canonical: particles filed in global-id order before the stable sort (the composite (cell key, id) order of SURVEY.md
    8e) -- the two-rank model must then equal the single-domain oracle BIT FOR BIT, not just within the tolerance."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "slab_model_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2", SLAB_MODEL_CANONICAL="1" if canonical else "0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SLAB_MODEL ok=True canonical=%s" % canonical in r.stdout


def test_rebalance_planes_moves_towards_equal_shares():
    from pbf_b200.slab import rebalance_planes
    counts = np.zeros(64, np.int64)
    counts[10:40] = 100                        # all the fluid in layers 10..39
    planes = [0, 32, 64]                       # rank 0 holds 2200, rank 1 holds 800
    for _ in range(10):
        new = rebalance_planes(counts, planes, max_shift=1)
        assert abs(new[1] - planes[1]) <= 1 and new[0] == 0 and new[2] == 64
        planes = new
    assert planes[1] == 25                     # equal shares: 1500 particles either side
    # thickness limit: the lower slab may not grow past what its window was allocated for
    assert rebalance_planes(counts, [0, 32, 64], max_shift=8, max_thickness=33)[1] >= 31
    # three ranks, minimum two layers each
    p3 = rebalance_planes(counts, [0, 2, 4, 64], max_shift=50)
    assert all(b - a >= 2 for a, b in zip(p3[:-1], p3[1:])) and p3[0] == 0 and p3[-1] == 64
