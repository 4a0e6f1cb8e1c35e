"""World-size-2 tests of the host-side multi-rank logic on the gloo backend (CPU only).

The GPU data path shards TARGET bodies over ranks and all-gathers positions once per step.  Here
the same partition / exchange layout drives the CPU oracle instead of the CUDA kernels, so the
test checks the launcher logic -- ranges tile the body array, the gathered layout is rank-ordered,
opaque ids reach every rank, timings reduce with MAX -- and that target sharding with a per-step
position all-gather reproduces the single-process result bit for bit."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, steps, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module("procedural-universe_b200")
    multi = importlib.import_module("procedural-universe_b200.multi")
    from oracle import port as oracle

    assert multi.env_rank_world() == (rank, world, rank)
    first, count = multi.shard_range(n, rank, world)
    # 1. opaque 128-byte id travels from rank 0
    payload = bytes(range(128)) if rank == 0 else None
    assert multi.broadcast_bytes(payload) == bytes(range(128))
    # 2. MAX over ranks
    assert multi.max_over_ranks(10.0 + rank) == 10.0 + world - 1
    # 3. sharded all-pairs steps with a per-step position all-gather
    p = pkg.seed_galaxy_host(n, 42, 1.0)
    dt = np.float32(0.01)
    for _ in range(steps):
        f = oracle.allpairs_forces(p, first, count)
        mine = p[first:first + count].copy()
        a = f / mine["Mass"][:, None]
        mine["Velocity"] += a * np.float64(dt)
        mine["Position"] += ((mine["Velocity"] * np.float64(dt)) / (20 * 1.15e12)).astype(np.float32)
        p = multi.all_gather_rows(mine, n, rank, world)          # rank-ordered concatenation
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), Position=p["Position"], Velocity=p["Velocity"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 512), (2, 301), (3, 100)])
def test_sharded_steps_equal_single_process(tmp_path, world, n):
    steps = 3
    mp.spawn(_worker, args=(world, _free_port(), n, steps, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("procedural-universe_b200")
    from oracle import port as oracle
    want = oracle.allpairs_run(pkg.seed_galaxy_host(n, 42, 1.0), 0.01, steps)
    for r in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{r}.npz"))
        assert np.array_equal(got["Position"], want["Position"])
        assert np.array_equal(got["Velocity"], want["Velocity"])


def test_shard_ranges_tile_the_array():
    sys.path.insert(0, ROOT)
    multi = importlib.import_module("procedural-universe_b200.multi")
    for n, world in [(1 << 20, 8), (1000, 3), (7, 7), (1 << 24, 8), (5, 2)]:
        edges = [multi.shard_range(n, r, world) for r in range(world)]
        assert edges[0][0] == 0 and sum(c for _, c in edges) == n
        for (f0, c0), (f1, _) in zip(edges, edges[1:]):
            assert f0 + c0 == f1
    with pytest.raises(Exception):
        multi.shard_range(10, 3, 3)
