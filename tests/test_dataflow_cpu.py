"""euler_b200/dataflow.py (UniqueDataFlow / SageDataFlow block construction) against a literal numpy restatement of
tf_euler/python/dataflow/{neighbor,sage}_dataflow.py, with a CPU stand-in sampler (oracle sampling + numpy first-occurrence
unique).  The device ops it composes (sample_neighbor, unique) have their own GPU parity tests."""
import numpy as np
import torch

import graphs
from oracle import pyoracle as po


def np_unique_first(x):
    vals, first, inv = np.unique(x, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")          # first-occurrence order (tf.unique)
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return vals[order], rank[inv]


class CpuSampler:
    def __init__(self, g, seed):
        self.og = graphs.oracle_graph(g)
        po.seed(seed)

    def sample_neighbor(self, nodes, edge_types, count, default_node=-1):
        ids, w, t = self.og.op_sample_neighbor(nodes.numpy().astype(np.int64), edge_types, count, default_node)
        return torch.from_numpy(ids.reshape(-1, count)), torch.from_numpy(w.reshape(-1, count)), torch.from_numpy(t.reshape(-1, count))

    def unique(self, ids):
        v, inv = np_unique_first(ids.numpy())
        return torch.from_numpy(v), torch.from_numpy(inv.astype(np.int32))


def reference_sage_flow(og, n_id, metapath, fanouts, max_id, add_self_loops):
    """literal restatement: sage_dataflow.py:35-50 then neighbor_dataflow.py:84-109"""
    neighbors, srcs = [], []
    cur = n_id.reshape(-1)
    for et, c in zip(metapath, fanouts):
        one, _, _ = og.op_sample_neighbor(cur, et, c, max_id + 1)
        one = one.reshape(-1)
        neighbors.append(one)
        srcs.append(np.repeat(np.arange(len(cur)), c))
        cur, _ = np_unique_first(np.concatenate([one, cur]))
    blocks = []
    cur = n_id.reshape(-1)
    last_idx = np.arange(len(cur))
    for i in range(len(metapath)):
        new = np.concatenate([neighbors[i], cur])
        new_u, inv = np_unique_first(new)
        res = inv[-len(cur):]
        src = srcs[i]
        if add_self_loops:
            src = np.concatenate([src, last_idx])
            last_idx = np.arange(len(new_u))
            dst = inv
        else:
            dst = inv[:-len(cur)]
            last_idx = dst
        blocks.append((new_u, res, np.stack([src, dst]), (len(cur), len(new_u))))
        cur = new_u
    return blocks[::-1]


def test_sage_dataflow_blocks_equal_the_reference_construction():
    from euler_b200.dataflow import SageDataFlow
    g = graphs.random_graph(seed=11, n=800, T=2, avg_deg=4, id_stride=3, id_base=2, hub=90)
    roots = g["ids"][np.random.RandomState(3).randint(0, 800, size=64)].astype(np.int64)
    roots[::9] = 10 ** 9          # absent ids -> default fill max_id + 1
    for self_loops in (True, False):
        sampler = CpuSampler(g, 77)
        flow = SageDataFlow([5, 3], [[0, 1], [1]], add_self_loops=self_loops, max_id=10 ** 7, sampler=sampler)(torch.from_numpy(roots))
        po.seed(77)
        want = reference_sage_flow(graphs.oracle_graph(g), roots, [[0, 1], [1]], [5, 3], 10 ** 7, self_loops)
        assert len(flow) == 2
        for blk, (n_id, res, ei, size) in zip(flow, want):
            assert np.array_equal(blk.n_id.numpy(), n_id)
            assert np.array_equal(blk.res_n_id.numpy(), res)
            assert np.array_equal(blk.edge_index.numpy(), ei)
            assert blk.size == size
            # every edge's source value is the sampled neighbor / self loop it claims
            assert ei[1].max() < len(n_id) and ei[0].max() < size[0]
