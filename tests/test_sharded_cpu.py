"""CPU, world_size 2 over gloo: the sharded orchestration (euler_b200/sharded.py -- routing, exchange
order, merge, frontier chaining, feature fetch) with oracle-backed per-shard ops, against a
single-process restatement of the sharded semantics."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases
import graphs
import sharded_common as sc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from euler_b200.sharded import ShardedGraph, TorchExchange
        g = graphs.random_graph(seed=77, n=600, T=2, avg_deg=5, feat_dim=6, id_stride=3, id_base=4, hub=80)
        shards = sc.partition(g, world)
        rs = [np.random.RandomState(100 + r) for r in range(world)]
        seeds = [g["ids"][x.randint(0, 600, size=150)].astype(np.int64) for x in rs]
        for s in seeds:
            s[::11] = 999983  # absent id
            s[3::13] = 0
        ets, counts = [[0, 1], [1, 0]], [4, 3]
        expect = sc.simulate(shards, seeds, ets, counts, shard_seeds=[500 + s for s in range(world)])
        ops = sc.OracleShardOps(shards[rank], 500 + rank)
        sg = ShardedGraph(ops, TorchExchange())
        ids, ws, ts = sg.sample_fanout(seeds[rank], ets, counts, -1)
        for l in range(3):
            cases.eq(ids[l].numpy(), expect[rank][0][l], "rank %d ids hop %d" % (rank, l))
        for l in range(2):
            cases.eq(ws[l].numpy(), expect[rank][1][l], "rank %d w hop %d" % (rank, l))
            cases.eq(ts[l].numpy(), expect[rank][2][l], "rank %d t hop %d" % (rank, l))
        # features of everything sampled == the unsharded oracle's features (no randomness involved)
        full = graphs.oracle_graph(g)
        f = sg.get_dense_feature(ids[2], 0, 6)
        cases.eq(f.numpy(), full.op_get_dense_feature(ids[2].numpy(), 6), "rank %d features" % rank)
        # every sampled edge is an edge of the unsharded graph
        src = np.repeat(ids[1].numpy(), 3)
        for a, b in list(zip(src, ids[2].numpy()))[::17]:
            if b != -1:
                lens, nb, _, _ = full.get_full_neighbor([a], [0, 1])
                assert b in nb.tolist()
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


def test_sharded_fanout_and_features_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    for r, msg in got:
        assert msg == "ok", "rank %d:\n%s" % (r, msg)


def test_partition_is_a_partition():
    g = graphs.random_graph(seed=5, n=100, T=3)
    shards = sc.partition(g, 4, P=8)
    assert sorted(np.concatenate([s["ids"] for s in shards]).tolist()) == g["ids"].tolist()
    assert sum(len(s["nbr"]) for s in shards) == len(g["nbr"])
    from euler_b200.sharded import owner_of
    for k, s in enumerate(shards):
        assert (owner_of(s["ids"], 8, 4) == k).all()
