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
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("SHARDED_GPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
