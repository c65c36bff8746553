"""Multi-GPU path: the graph is hash-partitioned by node id over the ranks of one box (Euler's own
shard scheme) and every hop / feature fetch resolves remote ids with an all-to-all over NVLink.

Reference being replaced (mode=remote): ID_SPLIT -> REMOTE (gRPC Execute of a sub-DAG on each shard
server) -> IDX_MERGE / DATA_MERGE, euler/core/kernels/id_split_op.cc:46-99, remote_op.cc:60-146,
idx_merge_op.cc:32-78, discovery through ZooKeeper.  Here: one process per GPU, torch.distributed
(NCCL) for the exchange, CUDA kernels (csrc/shard.cu) for routing and merging.

Per hop on every rank
    bucket seeds by owner (stable)            eu_shard_bucket
    exchange per-owner counts                 all_to_all (N ints; the only host sync of the hop)
    exchange seeds                            all_to_all_v  (8 B / seed)
    sample the received seeds on this shard   eu_sample_neighbor  (this shard's own engine)
    exchange the rows back                    all_to_all_v  (one 16 B record {id, w | t<<32} per slot)
    merge into request order + TF packing     eu_shard_merge_sample

Determinism / parity: a shard processes the concatenation of the requests of rank 0, 1, ... (each in
batch order) as ONE sampleNB call on its own engine, so duplicate seeds -- also across requesters --
share one sample row and the draw order is fixed.  The reference runs every request on whichever
server thread picks it up (thread_local engines, random.cc:22), i.e. it has no defined order here;
tests/test_sharded_*.py pin this definition against the oracle.

The exchange and the per-shard ops are injected (`ops`, `xchg`) so the same orchestration runs over
gloo on CPU in the tests.
"""
import ctypes as C

import numpy as np


def owner_of(ids, num_partitions, shard_num, self_shard=None):
    """euler/core/kernels/id_split_op.cc:46-49 (ids as uint64).  With self_shard given, ids 0 and 2^64-1
    (engine placeholder / default fill: they exist on no shard) stay on the requesting rank."""
    a = np.asarray(ids).astype(np.uint64)
    own = (a % np.uint64(num_partitions)) % np.uint64(shard_num)
    if self_shard is not None:
        own = np.where((a == 0) | (a == np.uint64(0xFFFFFFFFFFFFFFFF)), np.uint64(self_shard), own)
    return own


class TorchExchange:
    """all-to-all over a torch.distributed process group (NCCL for CUDA tensors, gloo for CPU)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def counts(self, send_counts_t):
        """send_counts_t: int64[world] tensor (device of the backend).  Returns (send, recv) lists."""
        import torch
        recv = torch.empty_like(send_counts_t)
        self.dist.all_to_all_single(recv, send_counts_t, group=self.group)
        both = torch.stack([send_counts_t, recv]).cpu()   # the hop's single host sync
        return both[0].tolist(), both[1].tolist()

    def a2a(self, t, send, recv, width=1):
        """t: [sum(send) * width] flat tensor; returns [sum(recv) * width]."""
        import torch
        out = torch.empty(int(sum(recv)) * width, dtype=t.dtype, device=t.device)
        self.dist.all_to_all_single(out, t, [int(r) * width for r in recv], [int(s) * width for s in send],
                                    group=self.group)
        return out


class CudaShardOps:
    """Per-shard work on this rank's GPU through the C ABI."""

    def __init__(self, graph, rng="minstd", seed=1):
        import torch
        from . import _lib
        from .graph import Context
        self.torch, self.lib, self.check = torch, _lib.load(), _lib.check
        self.graph = graph
        self.dev = torch.device("cuda", graph.device)
        self.ctx = Context(graph, rng, seed)

    def _stream(self):
        self.ctx.set_stream(self.torch.cuda.current_stream(self.dev).cuda_stream)
        return self.ctx._h

    def to_dev(self, a, dtype):
        t = self.torch
        if isinstance(a, t.Tensor):
            return a.to(device=self.dev, dtype=dtype).contiguous()
        return t.as_tensor(np.asarray(a), dtype=dtype, device=self.dev).contiguous()

    def bucket(self, ids, P, N, me):
        t = self.torch
        rows = ids.numel()
        sorted_ids = t.empty(rows, dtype=t.int64, device=self.dev)
        src = t.empty(rows, dtype=t.int32, device=self.dev)
        counts = t.empty(N, dtype=t.int64, device=self.dev)
        offs = t.empty(N + 1, dtype=t.int64, device=self.dev)
        self.check(self.lib.eu_shard_bucket(self._stream(), ids.data_ptr(), rows, P, N, me, sorted_ids.data_ptr(),
                                            src.data_ptr(), counts.data_ptr(), offs.data_ptr()))
        return sorted_ids, src, counts

    def sample_local(self, seeds, etypes, count):
        """Engine-form rows for the received seeds: ids with 0 placeholders, w 0, t -1 on empty rows."""
        t = self.torch
        n = seeds.numel()
        et = np.ascontiguousarray(etypes, dtype=np.int32)
        ids = t.empty(n * count, dtype=t.int64, device=self.dev)
        w = t.empty(n * count, dtype=t.float32, device=self.dev)
        ty = t.empty(n * count, dtype=t.int32, device=self.dev)
        self.check(self.lib.eu_sample_neighbor(self._stream(), seeds.data_ptr(), n, et.ctypes.data, len(et), count, 0,
                                               ids.data_ptr(), w.data_ptr(), ty.data_ptr()))
        packed = t.empty(n * count * 2, dtype=t.int64, device=self.dev)
        self.check(self.lib.eu_shard_pack_sample(self._stream(), ids.data_ptr(), w.data_ptr(), ty.data_ptr(), n * count,
                                                 packed.data_ptr()))
        return packed

    def merge_sample(self, packed, src, rows, count, default_node):
        t = self.torch
        eng = t.empty(rows * count, dtype=t.int64, device=self.dev)
        o_ids = t.empty(rows * count, dtype=t.int64, device=self.dev)
        o_w = t.empty(rows * count, dtype=t.float32, device=self.dev)
        o_t = t.empty(rows * count, dtype=t.int32, device=self.dev)
        self.check(self.lib.eu_shard_merge_sample(self._stream(), packed.data_ptr(),
                                                  src.data_ptr(), rows, count, default_node, eng.data_ptr(),
                                                  o_ids.data_ptr(), o_w.data_ptr(), o_t.data_ptr()))
        return eng, o_ids, o_w, o_t

    def feature_local(self, ids, fid, dim):
        t = self.torch
        out = t.empty(ids.numel() * dim, dtype=t.float32, device=self.dev)
        self.check(self.lib.eu_get_dense_feature(self._stream(), ids.data_ptr(), ids.numel(), fid, dim, out.data_ptr()))
        return out

    def merge_rows(self, rows_in, src, rows, dim):
        t = self.torch
        out = t.empty((rows, dim), dtype=t.float32, device=self.dev)
        self.check(self.lib.eu_shard_merge_rows(self._stream(), rows_in.data_ptr(), src.data_ptr(), rows, dim,
                                                out.data_ptr()))
        return out

    def seed(self, s):
        self.ctx.seed(s)


class ShardedGraph:
    """tf_euler's sampling / feature ops over a graph partitioned across the ranks of `xchg`."""

    def __init__(self, ops, xchg, num_partitions=None):
        self.ops, self.xchg = ops, xchg
        self.N = xchg.world
        self.P = num_partitions or self.N   # partitions a multiple of shards -> owner = id % N

    def sample_neighbor(self, nodes, edge_types, count, default_node=-1):
        eng, ids, w, t = self._hop(self.ops.to_dev(nodes, _i64(self.ops)), edge_types, count, default_node)
        shape = (-1, count)
        return ids.reshape(shape), w.reshape(shape), t.reshape(shape)

    def sample_fanout(self, nodes, edge_types, counts, default_node=-1):
        """neighbor_ops.sample_fanout (tf_euler/python/euler_ops/neighbor_ops.py:122-158): the next hop's
        seeds are the ENGINE ids of this hop (0 placeholders), tf_euler/kernels/sample_fanout_op.cc:36-43."""
        frontier = self.ops.to_dev(nodes, _i64(self.ops)).reshape(-1)
        ids, ws, ts = [frontier], [], []
        for et, c in zip(edge_types, counts):
            frontier, o_ids, o_w, o_t = self._hop(frontier, et, int(c), default_node)
            ids.append(o_ids); ws.append(o_w); ts.append(o_t)
        return ids, ws, ts

    def _hop(self, frontier, etypes, count, default_node):
        ops, x = self.ops, self.xchg
        rows = frontier.numel()
        sorted_ids, src, counts = ops.bucket(frontier, self.P, self.N, x.rank)
        send, recv = x.counts(counts)
        inbox = x.a2a(sorted_ids, send, recv)
        packed = ops.sample_local(inbox, etypes, count)            # 16-byte records {id, w | t << 32}
        back = x.a2a(packed, recv, send, count * 2)
        return ops.merge_sample(back, src, rows, count, default_node)

    def get_dense_feature(self, nodes, fid, dim):
        ops, x = self.ops, self.xchg
        ids = ops.to_dev(nodes, _i64(ops)).reshape(-1)
        rows = ids.numel()
        sorted_ids, src, counts = ops.bucket(ids, self.P, self.N, x.rank)
        send, recv = x.counts(counts)
        inbox = x.a2a(sorted_ids, send, recv)
        feats = ops.feature_local(inbox, fid, dim)
        back = x.a2a(feats, recv, send, dim)
        return ops.merge_rows(back, src, rows, dim)


class PeerShardedGraph:
    """The same ops as ShardedGraph with the exchange done by the kernels themselves over NVLink peer memory
    (csrc/p2p.cu): no NCCL call and no host sync per hop, so a whole step can be captured in a CUDA graph.
    torch.distributed is used once, at construction, to all_gather the cudaIpc handles."""

    def __init__(self, graph, rank, world, max_rows, max_count, max_feat_rows, max_dim, rng="minstd", seed=1,
                 num_partitions=None, group=None, engines=1):
        """max_rows / max_feat_rows count ALL batches of a batched call; engines = the largest nb used."""
        import torch
        import torch.distributed as dist
        from . import _lib
        from .graph import Context
        self.torch, self.lib, self.check = torch, _lib.load(), _lib.check
        self.graph, self.rank, self.N = graph, rank, world
        self.P = num_partitions or world
        self.dev = torch.device("cuda", graph.device)
        self.ctx = Context(graph, rng, seed)
        if engines > 1:
            self.ctx.set_engines(engines)
        self.ctx.reserve(world * max_rows + 1024)
        self._h = C.c_void_p()
        handle = (C.c_char * 64)()
        self.check(self.lib.eu_sym_create(self.ctx._h, rank, world, max_rows, max_count, max_feat_rows, max_dim,
                                          C.byref(self._h), handle))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        blob = b"".join(handles)
        self.check(self.lib.eu_sym_connect(self._h, blob))
        dist.barrier(group=group)
        ptrs = [C.c_void_p() for _ in range(5)]
        self.check(self.lib.eu_sym_outputs(self._h, *[C.byref(p) for p in ptrs]))
        mk = self._view
        n_out = max_rows * max_count
        self.o_eng = mk(ptrs[0].value, n_out, torch.int64)
        self.o_ids = mk(ptrs[1].value, n_out, torch.int64)
        self.o_w = mk(ptrs[2].value, n_out, torch.float32)
        self.o_t = mk(ptrs[3].value, n_out, torch.int32)
        self.o_rows = mk(ptrs[4].value, max_feat_rows * max_dim, torch.float32)

    def _view(self, ptr, n, dtype):
        """zero-copy tensor over library-owned device memory"""
        t = self.torch
        nbytes = max(n, 1) * t.empty(0, dtype=dtype).element_size()
        st = t._C._construct_storage_from_data_pointer(ptr, self.dev, nbytes)
        return t.empty(0, dtype=dtype, device=self.dev).set_(st, 0, (max(n, 1),))

    def _stream(self):
        self.ctx.set_stream(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def error(self):
        e = C.c_int(0)
        self.check(self.lib.eu_sym_error(self._h, C.byref(e)))
        return e.value

    def hop(self, frontier, etypes, count, default_node=-1, packed=True, nb=1):
        """frontier: device i64 tensor, [nb * rows] (batch-major).  Returns (eng, ids, w, t) VIEWS into the symmetric
        outputs (overwritten by the next hop; eng is what the next hop consumes), each [nb * rows * count]."""
        self._stream()
        et = np.ascontiguousarray(etypes, dtype=np.int32)
        total = frontier.numel()
        assert total % nb == 0
        self.check(self.lib.eu_sym_sample_hop_batched(self._h, frontier.data_ptr(), int(nb), total // nb, et.ctypes.data, len(et),
                                                      int(count), default_node, self.P, int(packed)))
        n = total * int(count)
        return self.o_eng[:n], self.o_ids[:n], self.o_w[:n], self.o_t[:n]

    def sample_fanout_batched(self, nodes, edge_types, counts, default_node=-1):
        """nodes: [nb, B].  Batch g == sample_fanout of nodes[g] with every shard on its engine g.  Returns lists of
        [nb, B * prod(counts[:l])] tensors."""
        t = self.torch
        nodes = nodes if isinstance(nodes, t.Tensor) else t.as_tensor(np.asarray(nodes), dtype=t.int64)
        nodes = nodes.to(device=self.dev, dtype=t.int64).contiguous()
        nb = nodes.shape[0]
        frontier = nodes.reshape(-1)
        ids, ws, ts = [nodes], [], []
        for et, c in zip(edge_types, counts):
            eng, o_ids, o_w, o_t = self.hop(frontier, et, c, default_node, nb=nb)
            ids.append(o_ids.clone().reshape(nb, -1)); ws.append(o_w.clone().reshape(nb, -1)); ts.append(o_t.clone().reshape(nb, -1))
            frontier = eng.clone()
        return ids, ws, ts

    def sample_fanout(self, nodes, edge_types, counts, default_node=-1):
        t = self.torch
        frontier = nodes if isinstance(nodes, t.Tensor) else t.as_tensor(np.asarray(nodes), dtype=t.int64, device=self.dev)
        frontier = frontier.to(device=self.dev, dtype=t.int64).reshape(-1).contiguous()
        ids, ws, ts = [frontier], [], []
        for et, c in zip(edge_types, counts):
            eng, o_ids, o_w, o_t = self.hop(frontier, et, c, default_node)
            ids.append(o_ids.clone()); ws.append(o_w.clone()); ts.append(o_t.clone())
            frontier = eng.clone()
        return ids, ws, ts

    def get_dense_feature(self, nodes, fid, dim, clone=True):
        t = self.torch
        ids = nodes.to(device=self.dev, dtype=t.int64).reshape(-1).contiguous()
        self._stream()
        self.check(self.lib.eu_sym_get_dense_feature(self._h, ids.data_ptr(), ids.numel(), int(fid), int(dim), self.P))
        out = self.o_rows[:ids.numel() * dim].reshape(ids.numel(), dim)
        return out.clone() if clone else out

    def sage_mean(self, nbr_ids, rows, count, dim, out=None):
        """fused sharded mean aggregation of a fixed-fanout block (owners sum their rows; eu_sym_sage_mean)"""
        t = self.torch
        ids = nbr_ids.to(device=self.dev, dtype=t.int64).reshape(-1).contiguous()
        assert ids.numel() == rows * count
        if out is None:
            out = t.empty(rows, dim, dtype=t.float32, device=self.dev)
        self._stream()
        self.check(self.lib.eu_sym_sage_mean(self._h, ids.data_ptr(), int(rows), int(count), int(dim), self.P, out.data_ptr()))
        return out

    def close(self):
        if self._h:
            self.torch.cuda.synchronize()
            self.lib.eu_sym_destroy(self._h)
            self._h = None


def _i64(ops):
    return ops.torch.int64 if hasattr(ops, "torch") else np.int64
