"""Launched by torchrun on >= 2 GPUs (tests/test_sharded_gpu.py or by hand):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/run_sharded_gpu.py
Each rank holds its shard on its own GPU; the CUDA sharded fanout / feature fetch must equal the
single-process oracle restatement of the sharded semantics (tests/sharded_common.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import graphs  # noqa: E402
import sharded_common as sc  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from euler_b200.sharded import CudaShardOps, ShardedGraph, TorchExchange
    g = graphs.random_graph(seed=91, n=30000, T=3, avg_deg=7, feat_dim=64, id_stride=5, id_base=2, hub=3000, zero_w_frac=0.05)
    shards = sc.partition(g, world)
    rs = [np.random.RandomState(300 + r) for r in range(world)]
    seeds = [g["ids"][x.randint(0, 30000, size=2048)].astype(np.int64) for x in rs]
    for s in seeds:
        s[::11] = 999999937
        s[3::13] = 0
        s[5::17] = -1
    ets, counts = [[0, 2], [2, 1]], [25, 10]
    expect = sc.simulate(shards, seeds, ets, counts, shard_seeds=[700 + s for s in range(world)])
    gr = graphs.cuda_graph(shards[rank], device=local)
    ops = CudaShardOps(gr, "minstd", 700 + rank)
    sg = ShardedGraph(ops, TorchExchange())
    ids, ws, ts = sg.sample_fanout(seeds[rank], ets, counts, -1)
    for l in range(3):
        cases.eq(ids[l].cpu().numpy(), expect[rank][0][l], "rank %d ids hop %d" % (rank, l))
    for l in range(2):
        cases.eq(ws[l].cpu().numpy(), expect[rank][1][l], "rank %d w hop %d" % (rank, l))
        cases.eq(ts[l].cpu().numpy(), expect[rank][2][l], "rank %d t hop %d" % (rank, l))
    full = graphs.oracle_graph(g)
    f = sg.get_dense_feature(ids[2], 0, 64)
    cases.eq(f.cpu().numpy(), full.op_get_dense_feature(ids[2].cpu().numpy(), 64), "rank %d features" % rank)
    # second call continues every shard's engine stream
    expect2 = None
    ids2, _, _ = sg.sample_fanout(seeds[rank], ets, counts, -1)
    assert not np.array_equal(ids2[1].cpu().numpy(), ids[1].cpu().numpy())
    # ---- the same semantics with the exchange done by the kernels over NVLink peer memory (csrc/p2p.cu)
    from euler_b200.sharded import PeerShardedGraph
    n1, n2 = 2048 * 25, 2048 * 250
    pg = PeerShardedGraph(gr, rank, world, max_rows=n1, max_count=25, max_feat_rows=n2, max_dim=64, rng="minstd", seed=700 + rank)
    p_ids, p_ws, p_ts = pg.sample_fanout(seeds[rank], ets, counts, -1)
    assert pg.error() == 0, "peer exchange timed out"
    for l in range(3):
        cases.eq(p_ids[l].cpu().numpy(), expect[rank][0][l], "peer: rank %d ids hop %d" % (rank, l))
    for l in range(2):
        cases.eq(p_ws[l].cpu().numpy(), expect[rank][1][l], "peer: rank %d w hop %d" % (rank, l))
        cases.eq(p_ts[l].cpu().numpy(), expect[rank][2][l], "peer: rank %d t hop %d" % (rank, l))
    pf = pg.get_dense_feature(p_ids[2], 0, 64)
    cases.eq(pf.cpu().numpy(), full.op_get_dense_feature(p_ids[2].cpu().numpy(), 64), "peer: rank %d features" % rank)
    for _ in range(3):   # repeated exchanges reuse the inboxes / epochs
        q_ids, _, _ = pg.sample_fanout(seeds[rank], ets, counts, -1)
        qf = pg.get_dense_feature(q_ids[1], 0, 64)
    assert pg.error() == 0
    cases.eq(qf.cpu().numpy(), full.op_get_dense_feature(q_ids[1].cpu().numpy(), 64), "peer: features after reuse")
    # ---- fused sharded SAGE mean: owners sum their rows, requester adds the partials in rank order (bit-exact association)
    agg = pg.sage_mean(q_ids[1], 2048, 25, 64)              # generic-width kernel (dim 64)
    cases.eq(agg.cpu().numpy(), sc.sage_mean_sharded(full, q_ids[1].cpu().numpy(), 2048, 25, 64, world, rank), "peer: sage mean dim 64")
    agg2 = pg.sage_mean(q_ids[2][:4096 * 10], 4096, 10, 64)
    cases.eq(agg2.cpu().numpy(), sc.sage_mean_sharded(full, q_ids[2][:4096 * 10].cpu().numpy(), 4096, 10, 64, world, rank), "peer: sage mean hop 2")
    assert pg.error() == 0
    torch.cuda.synchronize()
    dist.barrier()
    pg.close()
    # ---- nb batches per exchange: batch b == the sharded fanout with every shard on its engine b (seed + b)
    NB = 3
    bs = [[g["ids"][np.random.RandomState(900 + 10 * r + b).randint(0, 30000, size=512)].astype(np.int64) for b in range(NB)] for r in range(world)]
    for r in range(world):
        bs[r][1][::5] = 0
        bs[r][2][::9] = 31337
        bs[r][0][:64] = bs[r][0][64:128]          # duplicates inside a batch
    pgb = PeerShardedGraph(gr, rank, world, max_rows=NB * 512 * 25, max_count=25, max_feat_rows=NB * 512 * 25, max_dim=64, rng="minstd",
                           seed=4100 + 100 * rank, engines=NB)
    for rep in range(2):
        b_ids, b_ws, b_ts = pgb.sample_fanout_batched(np.stack(bs[rank]), ets, counts, -1)
    assert pgb.error() == 0
    # expectation: per batch, an independent simulation whose shard engines are seeded seed_s + b and called twice
    for b in range(NB):
        ops_seeds = [4100 + 100 * s_ + b for s_ in range(world)]
        exp_b = sc.simulate(shards, [bs[r][b] for r in range(world)], ets, counts, shard_seeds=ops_seeds, repeat=2)
        for l in range(3):
            cases.eq(b_ids[l][b].cpu().numpy(), exp_b[rank][0][l], "batched peer: rank %d batch %d ids hop %d" % (rank, b, l))
        for l in range(2):
            cases.eq(b_ws[l][b].cpu().numpy(), exp_b[rank][1][l], "batched peer: rank %d batch %d w hop %d" % (rank, b, l))
            cases.eq(b_ts[l][b].cpu().numpy(), exp_b[rank][2][l], "batched peer: rank %d batch %d t hop %d" % (rank, b, l))
    torch.cuda.synchronize()
    dist.barrier()
    pgb.close()
    # dim 128: the float4 warp kernel, fanout 40 > 32 (two lookup rounds per destination)
    g2 = graphs.random_graph(seed=92, n=8000, T=1, avg_deg=9, feat_dim=128, id_stride=3, id_base=1)
    sh2 = sc.partition(g2, world)
    gr2 = graphs.cuda_graph(sh2[rank], device=local)
    full2 = graphs.oracle_graph(g2)
    pg2 = PeerShardedGraph(gr2, rank, world, max_rows=1024, max_count=40, max_feat_rows=1024 * 40, max_dim=128, rng="minstd", seed=5 + rank)
    sd = g2["ids"][np.random.RandomState(40 + rank).randint(0, 8000, size=1024)].astype(np.int64)
    sd[::7] = 12345678901
    r_ids, _, _ = pg2.sample_fanout(sd, [[0]], [40], -1)
    for _ in range(2):
        agg3 = pg2.sage_mean(r_ids[1], 1024, 40, 128)
    want = sc.sage_mean_sharded(full2, r_ids[1].cpu().numpy(), 1024, 40, 128, world, rank)
    cases.eq(agg3.cpu().numpy(), want, "peer: sage mean dim 128")
    plain = full2.op_get_dense_feature(r_ids[1].cpu().numpy(), 128).reshape(1024, 40, 128).astype(np.float64).sum(1) / (40 + 1e-7)
    assert np.allclose(agg3.cpu().numpy(), plain, rtol=1e-5, atol=1e-6), "peer: sage mean vs f64 mean"
    assert pg2.error() == 0
    torch.cuda.synchronize()
    dist.barrier()
    pg2.close()
    # ---- replicated feature table (PeerShardedGraph(feature_graph=...)): the fetch and the fused mean are the single-GPU
    # kernels on a graph that holds every node's row -> bit-identical to the unsharded ops / the oracle's scatter_mean
    full_gr = graphs.cuda_graph(g, device=local)
    pr = PeerShardedGraph(gr, rank, world, max_rows=n1, max_count=25, max_feat_rows=1, max_dim=4, rng="minstd", seed=700 + rank,
                          feature_graph=full_gr)
    r_ids, r_ws, r_ts = pr.sample_fanout(seeds[rank], ets, counts, -1)
    for l in range(3):
        cases.eq(r_ids[l].cpu().numpy(), expect[rank][0][l], "replicated: rank %d ids hop %d" % (rank, l))
    rf = pr.get_dense_feature(r_ids[1], 0, 64)
    cases.eq(rf.cpu().numpy(), full.op_get_dense_feature(r_ids[1].cpu().numpy(), 64), "replicated: features")
    ra = pr.sage_mean(r_ids[2], n1, 10, 64)
    from oracle import pyoracle as po
    want_a = po.scatter_mean(full.op_get_dense_feature(r_ids[2].cpu().numpy(), 64), np.repeat(np.arange(n1, dtype=np.int32), 10), n1)
    cases.eq(ra.cpu().numpy(), want_a, "replicated: fused sage mean == oracle scatter_mean (bit-exact)")
    torch.cuda.synchronize()
    dist.barrier()
    pr.close()
    # ---- a rank that issues a different batch shape poisons the exchange on EVERY rank instead of corrupting it silently
    pz = PeerShardedGraph(gr, rank, world, max_rows=4096, max_count=4, max_feat_rows=1, max_dim=4, rng="minstd", seed=1 + rank)
    odd = seeds[rank][:1000 if rank == 0 else 1001]
    try:
        pz.sample_fanout(odd, [[0]], [4], -1)
        poisoned = False
    except Exception:
        poisoned = True
    assert poisoned, "rank %d: a batch-shape mismatch went undetected" % rank
    torch.cuda.synchronize()
    dist.barrier()
    try:
        pz.close()
    except Exception:
        pass
    # ---- sharded DeepWalk (p = q = 1) and sample_node on the CUDA per-shard ops over NCCL (host logic pinned over gloo in
    # tests/test_sharded_cpu.py with the oracle ops; here the same expectations with the real kernels)
    from euler_b200.sharded import ClientRng
    ops_w = CudaShardOps(gr, "minstd", 700 + rank)
    wseeds = [s_[:150] for s_ in seeds]
    walk = ShardedGraph(ops_w, TorchExchange()).random_walk(wseeds[rank], [[0, 1]] * 5, 1.0, 1.0, -1)
    exp_w = sc.simulate(shards, wseeds, [[0, 1]] * 5, [1] * 5, shard_seeds=[700 + s_ for s_ in range(world)])
    assert tuple(walk.shape) == (150, 6)
    for l in range(6):
        cases.eq(walk[:, l].cpu().numpy(), exp_w[rank][0][l], "cuda sharded walk: rank %d column %d" % (rank, l))
    g3 = graphs.random_graph(seed=78, n=900, T=1, avg_deg=3, n_node_types=3, id_stride=2, id_base=5)
    shards3 = sc.partition(g3, world)
    gr3 = graphs.cuda_graph(shards3[rank], device=local)
    for types in ([0], [-1], [1, 2]):
        ops3 = CudaShardOps(gr3, "minstd", 600 + rank)
        sg3 = ShardedGraph(ops3, TorchExchange())
        crng = ClientRng(900 + rank)
        for _ in range(2):
            got = sg3.sample_node(257, types, crng)
        want = sc.simulate_sample_node(shards3, 257, types, [600 + s_ for s_ in range(world)], [900 + r for r in range(world)], repeat=2)
        cases.eq(got.cpu().numpy(), want[rank], "cuda sharded sample_node: rank %d types %s" % (rank, types))
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("SHARDED_GPU_OK world=%d (peer exchange, batched, fused sage, replicated features, poison on shape mismatch, "
              "sharded walk + sample_node on CUDA ops)" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
