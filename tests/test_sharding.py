"""CPU suite: the N>1 host logic (query sharding, packed argmax all-reduce, result
concatenation) on world_size-2 and -3 gloo groups."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from my_lidar_graph_slam_v2_b200 import sharding


def test_shard_ranges_cover_queries_in_order():
    for nq in (1, 7, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            ranges = [sharding.shard_range(nq, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nq
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
    assert sharding.shard_range(256, 3, 8) == (96, 128)


def test_pack_orders_by_key_then_lowest_query():
    a, b = sharding.pack_best(1000, 5), sharding.pack_best(1001, 900)
    assert b > a
    assert sharding.pack_best(1000, 4) > sharding.pack_best(1000, 5)
    assert sharding.unpack_best(sharding.pack_best(19209203456, 130)) == (19209203456, 130)
    assert sharding.unpack_best(0) == (0, -1)
    # 1080 beams of value 65535: the key still fits next to the 20 index bits
    assert sharding.pack_best(998 * 1080 * 65535 + 64536 * 1080, sharding.QUERY_MASK) < (1 << 63)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nq, seed, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        keys = rng.integers(1, 10 ** 10, size=nq)
        keys[nq // 3] = keys.max()                      # force a tie between two queries
        keys[2 * nq // 3] = keys.max()
        found = rng.random(nq) < 0.4
        found[[nq // 3, 2 * nq // 3]] = True
        lo, hi = sharding.shard_range(nq, rank, world)
        word = torch.tensor([sharding.local_best_word(keys[lo:hi], found[lo:hi], lo)], dtype=torch.int64)
        sharding.allreduce_best(word)
        rec = np.stack([np.arange(lo, hi), keys[lo:hi], found[lo:hi].astype(np.int64)], axis=1)
        full = sharding.all_gather_results(rec)
        if rank == 0:
            out.put((int(word.item()), full.tolist(), keys.tolist(), found.tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nq", [(2, 256), (3, 100), (2, 5)])
def test_argmax_allreduce_and_gather_gloo(world, nq):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nq, 99, out)) for r in range(world)]
    for p in procs:
        p.start()
    word, full, keys, found = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    keys, found = np.asarray(keys), np.asarray(found)
    exp_key = keys[found].max()
    exp_q = int(np.nonzero(found & (keys == exp_key))[0][0])      # lowest index among equal keys
    assert sharding.unpack_best(word) == (int(exp_key), exp_q)
    full = np.asarray(full)
    assert full.shape == (nq, 3)
    assert np.array_equal(full[:, 0], np.arange(nq))               # rank order == query order
    assert np.array_equal(full[:, 1], keys) and np.array_equal(full[:, 2], found.astype(np.int64))
