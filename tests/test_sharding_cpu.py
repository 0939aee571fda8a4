"""N>1 path on CPU: scenes are sharded across ranks with no data-path collective; only the timing
reduction (max over ranks) and a barrier use the process group.  world_size 2, gloo."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200.sharding import shard_scenes, reduce_max_ms  # noqa: E402


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_scenes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard_scenes(total_scenes, rank, world)
    dist.barrier()
    ms = reduce_max_ms(10.0 + 5.0 * rank)      # rank 1 is slower
    gathered = [None] * world
    dist.all_gather_object(gathered, list(mine))
    q.put((rank, list(mine), ms, gathered))
    dist.destroy_process_group()


def test_scene_sharding_two_ranks_gloo():
    world, total = 2, 11
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [r[1] for r in res]
    assert sorted(shards[0] + shards[1]) == list(range(total))            # a partition of the scenes
    assert abs(len(shards[0]) - len(shards[1])) <= 1                      # balanced
    assert all(abs(r[2] - 15.0) < 1e-6 for r in res)                      # time = max over ranks


def test_shard_scenes_properties():
    for total in (0, 1, 7, 64):
        for world in (1, 2, 3, 8):
            parts = [list(shard_scenes(total, r, world)) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(total))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
