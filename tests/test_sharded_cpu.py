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
        # ---- replicated feature table: same rows without an exchange
        full_ops = sc.OracleShardOps(g, 1)
        f2 = ShardedGraph(sc.OracleShardOps(shards[rank], 501 + rank), TorchExchange(), feature_ops=full_ops).get_dense_feature(ids[2], 0, 6)
        cases.eq(f2.numpy(), f.numpy(), "rank %d replicated features" % rank)
        # ---- DeepWalk (p = q = 1) = chained count-1 hops: one exchange per step
        ops_w = sc.OracleShardOps(shards[rank], 700 + rank)
        walk = ShardedGraph(ops_w, TorchExchange()).random_walk(seeds[rank], [[0, 1]] * 5, 1.0, 1.0, -1)
        exp_w = sc.simulate(shards, seeds, [[0, 1]] * 5, [1] * 5, shard_seeds=[700 + s for s in range(world)])
        assert tuple(walk.shape) == (150, 6)
        for l in range(6):
            cases.eq(walk[:, l].numpy(), exp_w[rank][0][l], "rank %d walk column %d" % (rank, l))
        # ---- sharded sample_node: split by shard weight sums, remainder by the client's engine, shard-order merge
        from euler_b200.sharded import ClientRng
        g2 = graphs.random_graph(seed=78, n=900, T=1, avg_deg=3, n_node_types=3, id_stride=2, id_base=5)
        shards2 = sc.partition(g2, world)
        for types in ([0], [-1], [1, 2]):
            ops2 = sc.OracleShardOps(shards2[rank], 600 + rank)
            sg2 = ShardedGraph(ops2, TorchExchange())
            crng = ClientRng(900 + rank)
            for _ in range(2):
                got = sg2.sample_node(257, types, crng)
            want = sc.simulate_sample_node(shards2, 257, types, [600 + s for s in range(world)], [900 + r for r in range(world)], repeat=2)
            cases.eq(got.numpy(), want[rank], "rank %d sample_node %s" % (rank, types))
            assert got.numel() == 257
            tt = {int(i): int(t) for i, t in zip(g2["ids"], g2["node_type"])}
            if types != [-1]:
                assert all(tt[int(i)] in types for i in got.numpy())
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


def _run_world(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    got = [q.get(timeout=180) for _ in ps]
    for p in ps:
        p.join(timeout=30)
    for r, msg in got:
        assert msg == "ok", "rank %d:\n%s" % (r, msg)


def test_sharded_fanout_and_features_world2_gloo():
    _run_world(2)


def test_sharded_fanout_and_features_world3_gloo():
    """a shard count that is not a power of two: owner = (id % P) % 3, uneven segments"""
    _run_world(3)


def test_partition_is_a_partition():
    g = graphs.random_graph(seed=5, n=100, T=3)
    shards = sc.partition(g, 4, P=8)
    assert sorted(np.concatenate([s["ids"] for s in shards]).tolist()) == g["ids"].tolist()
    assert sum(len(s["nbr"]) for s in shards) == len(g["nbr"])
    from euler_b200.sharded import owner_of
    for k, s in enumerate(shards):
        assert (owner_of(s["ids"], 8, 4) == k).all()


def test_client_rng_and_split_match_the_oracle_engine():
    """ClientRng == the oracle's restatement of the reference engine; split_sample_count follows
    sample_node_split_op.cc:58-85 (floor shares + remainder draws)."""
    from euler_b200.sharded import ClientRng, shard_weight_table, split_sample_count
    from oracle import pyoracle as po
    for seed in (1, 12345, 0, 2 ** 31 - 1, 1758564000):
        a, b = ClientRng(seed), po.Rng(seed)
        for _ in range(2000):
            assert a.uniform() == b.uniform()
    table = shard_weight_table([[3.0, 0.0, 1.5], [1.0, 0.0, 2.5], [0.0, 0.0, 4.0]])     # 3 shards x 3 types
    assert table.shape == (4, 4) and table[3][3] == 12.0 and table[0][3] == 4.0
    r = ClientRng(7)
    sp = split_sample_count(10, [0], table, r)
    assert sum(sp) == 10 and sp[2] == 0 and sp[0] >= 7 and sp[1] >= 2          # floor(7.5), floor(2.5) + 1 leftover
    assert split_sample_count(12, [-1], table, ClientRng(1)) == [4, 3, 4] or sum(split_sample_count(12, [-1], table, ClientRng(1))) == 12
    with pytest.raises(ValueError):
        split_sample_count(5, [1], table, ClientRng(1))                          # zero total weight (EULER_LOG(FATAL) in the reference)
    with pytest.raises(ValueError):
        split_sample_count(5, [-1, 0], table, ClientRng(1))
