"""Minibatch subgraph construction for the message-passing convolutions: the reference's NeighborDataFlow /
UniqueDataFlow / SageDataFlow (tf_euler/python/dataflow/{base,neighbor,sage}_dataflow.py) over this package's ops.

A DataFlow is the list of Blocks a convolution stack consumes, deepest hop first (base_dataflow.py:19-52):
    block.n_id        node ids of the block's source side (hop l+1 frontier [+ the destination nodes])
    block.res_n_id    positions of the destination nodes inside n_id
    block.edge_index  [2, E]: edge_index[0] = index into the destination nodes, edge_index[1] = index into n_id
    block.size        (number of destination nodes, number of source nodes)

Tensors stay on the device; the only host sync is the unique count of UniqueDataFlow (a shape, as in TF).
`sampler` is any object with sample_neighbor(nodes, edge_types, count, default_node) -> (ids[B, count], w, t) and
unique(ids) -> (values, inverse): euler_b200 itself on the GPU, or a CPU stand-in in the tests."""
import torch


class Block:
    def __init__(self, n_id, res_n_id, e_id, edge_index, size):
        self.n_id, self.res_n_id, self.e_id, self.edge_index, self.size = n_id, res_n_id, e_id, edge_index, size


class DataFlow:
    """base_dataflow.py:31-52"""

    def __init__(self, n_id):
        self.n_id = n_id
        self._last = n_id
        self.blocks = []

    def append(self, n_id, res_n_id, e_id, edge_index):
        self.blocks.append(Block(n_id, res_n_id, e_id, edge_index, (int(self._last.numel()), int(n_id.numel()))))
        self._last = n_id

    def __len__(self):
        return len(self.blocks)

    def __getitem__(self, idx):
        return self.blocks[::-1][idx]

    def __iter__(self):
        return iter(self.blocks[::-1])


def _default_sampler():
    import euler_b200
    return euler_b200


class NeighborDataFlow:
    """neighbor_dataflow.py:24-73: no de-duplication; block l's sources are [hop-l neighbors, destinations]."""

    def __init__(self, num_hops, add_self_loops=True, sampler=None):
        self.num_hops, self.add_self_loops = num_hops, add_self_loops
        self.sampler = sampler or _default_sampler()

    def get_neighbors(self, n_id):
        raise NotImplementedError

    def _merge(self, new_n_id):
        """identity numbering (no unique): values, inverse"""
        return new_n_id, torch.arange(new_n_id.numel(), device=new_n_id.device)

    def produce_subgraph(self, n_id):
        n_id = n_id.reshape(-1)
        last_idx = torch.arange(n_id.numel(), device=n_id.device)
        flow = DataFlow(n_id)
        neighbors, edge_srcs = self.get_neighbors(n_id)
        for i in range(self.num_hops):
            n_prev = n_id.numel()
            cat = torch.cat([neighbors[i], n_id])
            new_n_id, new_inv = self._merge(cat)
            res_n_id = new_inv[-n_prev:]
            edge_src = edge_srcs[i]
            if self.add_self_loops:
                edge_src = torch.cat([edge_src, last_idx])
                last_idx = torch.arange(new_n_id.numel(), device=n_id.device)
                edge_dst = new_inv
            else:
                edge_dst = new_inv[:-n_prev]
                last_idx = edge_dst
            n_id = new_n_id
            flow.append(new_n_id, res_n_id, None, torch.stack([edge_src.to(torch.int64), edge_dst.to(torch.int64)]))
        return flow

    __call__ = produce_subgraph


class UniqueDataFlow(NeighborDataFlow):
    """neighbor_dataflow.py:76-109: every block's source side is tf.unique'd (first-occurrence order)."""

    def _merge(self, new_n_id):
        values, inverse = self.sampler.unique(new_n_id)
        return values, inverse.to(torch.int64)


class SageDataFlow(UniqueDataFlow):
    """sage_dataflow.py:24-50: fixed-fanout sample_neighbor per hop; the next hop samples from unique(neighbors + nodes)."""

    def __init__(self, fanouts, metapath, add_self_loops=True, max_id=-1, sampler=None):
        super().__init__(num_hops=len(metapath), add_self_loops=add_self_loops, sampler=sampler)
        self.fanouts, self.metapath, self.max_id = fanouts, metapath, max_id

    def get_neighbors(self, n_id):
        neighbors, neighbor_src = [], []
        for hop_edge_types, count in zip(self.metapath, self.fanouts):
            n_id = n_id.reshape(-1)
            one, _w, _t = self.sampler.sample_neighbor(n_id, hop_edge_types, count, default_node=self.max_id + 1)
            one = one.reshape(-1)
            neighbors.append(one)
            neighbor_src.append(torch.arange(n_id.numel(), device=n_id.device).repeat_interleave(count))
            n_id, _ = self.sampler.unique(torch.cat([one, n_id]))
        return neighbors, neighbor_src
