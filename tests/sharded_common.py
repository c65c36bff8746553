"""Test infrastructure for the sharded path: graph partitioning, an oracle-backed per-shard ops
object (so euler_b200/sharded.py's orchestration can run over gloo on CPU), and a single-process
restatement of the sharded semantics to compare against.

Sharded semantics pinned here (see euler_b200/sharded.py): owner(id) = (id % P) % N
(euler/core/kernels/id_split_op.cc:46-49); per hop every shard processes the concatenation of the
requests of rank 0..N-1 (each in stable batch order, id_split_op.cc:70-75) as ONE engine sampleNB call
on its own engine stream."""
import numpy as np

import graphs
from oracle import pyoracle as po


def route(ids_u64, P, N, me):
    """owner of each id; 0 and 2^64-1 stay on the requesting rank (they exist nowhere)."""
    own = ((ids_u64 % np.uint64(P)) % np.uint64(N)).astype(np.int64)
    own[(ids_u64 == 0) | (ids_u64 == np.uint64(0xFFFFFFFFFFFFFFFF))] = me
    return own


def partition(g, N, P=None):
    """Split a tests/graphs.py graph dict into N shard graph dicts (rows keep their relative order)."""
    P = P or N
    own = (g["ids"] % np.uint64(P)) % np.uint64(N)
    T = g["T"]
    out = []
    for s in range(N):
        rows = np.nonzero(own == s)[0]
        ptr = [0]
        nbr, w, cum, gcum = [], [], [], []
        for r in rows:
            for t in range(T):
                b, e = g["grp_ptr"][r * T + t], g["grp_ptr"][r * T + t + 1]
                nbr.append(g["nbr"][b:e]); w.append(g["w"][b:e]); cum.append(g["cum_w"][b:e])
                ptr.append(ptr[-1] + (e - b))
            gcum.append(g["grp_cum"][r * T:(r + 1) * T])
        cat = lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dt)  # noqa: E731
        out.append(dict(ids=g["ids"][rows], node_type=g["node_type"][rows], node_w=g["node_w"][rows], T=T,
                        grp_ptr=np.asarray(ptr, np.int64), nbr=cat(nbr, np.uint64), w=cat(w, np.float32),
                        cum_w=cat(cum, np.float32), grp_cum=cat(gcum, np.float32),
                        feat=None if g.get("feat") is None else g["feat"][rows], n_node_types=g["n_node_types"]))
    return out


class OracleShardOps:
    """Per-shard ops on numpy / torch-CPU tensors, backed by the C oracle (its global engine = this
    shard's engine: one shard per process in the gloo tests)."""

    def __init__(self, shard_graph, seed):
        import torch
        self.torch = torch
        self.og = graphs.oracle_graph(shard_graph)
        self.shard = shard_graph
        self.og.build_node_sampler(np.arange(len(shard_graph["ids"]), dtype=np.int64), shard_graph["n_node_types"])
        self.feat_dim = 0 if shard_graph.get("feat") is None else shard_graph["feat"].shape[1]
        po.seed(seed)

    def to_dev(self, a, dtype):
        t = self.torch
        return (a if isinstance(a, t.Tensor) else t.as_tensor(np.asarray(a))).to(dtype).contiguous()

    def bucket(self, ids, P, N, me):
        t = self.torch
        a = ids.numpy().astype(np.uint64)
        own = route(a, P, N, me)
        order = np.argsort(own, kind="stable")
        counts = np.bincount(own, minlength=N).astype(np.int64)
        return t.from_numpy(a[order].astype(np.int64)), t.from_numpy(order.astype(np.int32)), t.from_numpy(counts)

    def sample_local(self, seeds, etypes, count):
        t = self.torch
        ids, w, ty = self.og.op_sample_neighbor(seeds.numpy(), etypes, count, 0)
        lo = w.reshape(-1).view(np.uint32).astype(np.uint64) | (ty.reshape(-1).astype(np.uint32).astype(np.uint64) << np.uint64(32))
        packed = np.stack([ids.reshape(-1), lo.view(np.int64)], axis=1).reshape(-1)
        return t.from_numpy(np.ascontiguousarray(packed))

    def merge_sample(self, packed, src, rows, count, default_node):
        t = self.torch
        rec = packed.numpy().reshape(rows, count, 2)
        r_ids = rec[:, :, 0]
        lo = rec[:, :, 1].view(np.uint64)
        r_w = (lo & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32)
        r_t = (lo >> np.uint64(32)).astype(np.uint32).view(np.int32)
        src = src.numpy()
        eng = np.zeros((rows, count), np.int64)
        eng[src] = r_ids
        keep = (r_ids[:, :1] != 0) if count else np.zeros((rows, 1), bool)
        o_ids = np.zeros((rows, count), np.int64); o_w = np.zeros((rows, count), np.float32); o_t = np.zeros((rows, count), np.int32)
        o_ids[src] = np.where(keep, r_ids, default_node)
        o_w[src] = np.where(keep, r_w, 0)
        o_t[src] = np.where(keep, r_t, -1)
        f = lambda x: t.from_numpy(x.reshape(-1))  # noqa: E731
        return f(eng), f(o_ids), f(o_w), f(o_t)

    def feature_local(self, ids, fid, dim):
        return self.torch.from_numpy(self.og.op_get_dense_feature(ids.numpy(), dim).reshape(-1))

    def node_weight_sums(self):
        return type_weight_sums(self.shard)

    def sample_node_local(self, n, node_types):
        """the shard's ONE engine (the oracle's global stream here) also serves its node draws"""
        st = po.get_state()
        r = po.Rng(1)
        r.x, r.draws = st[0], st[1]
        ids = self.og.sample_node(node_types, int(n), r) if n else np.zeros(0, np.uint64)
        po.set_state((r.x, r.draws))
        return self.torch.from_numpy(ids.astype(np.int64))

    def merge_rows(self, rows_in, src, rows, dim):
        out = np.zeros((rows, dim), np.float32)
        out[src.numpy()] = rows_in.numpy().reshape(rows, dim)
        return self.torch.from_numpy(out)


def simulate(shards, seeds_per_rank, ets, counts, shard_seeds, default_node=-1, P=None, repeat=1):
    """Single-process restatement: returns, per rank, (ids list per hop incl. hop 0, ws, ts) of the LAST of `repeat`
    identical calls (the shard engines keep running across calls)."""
    N = len(shards)
    P = P or N
    ogs = [graphs.oracle_graph(s) for s in shards]
    states = []
    for s in range(N):
        po.seed(shard_seeds[s])
        states.append(po.get_state())
    for _ in range(repeat):
        res = _simulate_once(ogs, states, seeds_per_rank, ets, counts, default_node, P, N)
    return res


def _simulate_once(ogs, states, seeds_per_rank, ets, counts, default_node, P, N):
    frontier = [np.asarray(x, np.int64) for x in seeds_per_rank]
    res = [([f.copy()], [], []) for f in frontier]
    for et, c in zip(ets, counts):
        own = [route(f.astype(np.uint64), P, N, r) for r, f in enumerate(frontier)]
        order = [np.argsort(o, kind="stable") for o in own]
        new_frontier = [np.zeros((len(f), c), np.int64) for f in frontier]
        packed = [[np.zeros((len(f), c), np.int64), np.zeros((len(f), c), np.float32), np.zeros((len(f), c), np.int32)] for f in frontier]
        for s in range(N):
            req_idx = [order[r][own[r][order[r]] == s] for r in range(N)]
            req = np.concatenate([frontier[r][req_idx[r]] for r in range(N)]) if N else np.zeros(0, np.int64)
            po.set_state(states[s])
            ids, w, t = ogs[s].op_sample_neighbor(req, et, c, 0)
            states[s] = po.get_state()
            off = 0
            for r in range(N):
                k = len(req_idx[r])
                ri, rw, rt = ids[off:off + k], w[off:off + k], t[off:off + k]
                off += k
                keep = ri[:, :1] != 0 if c else np.zeros((k, 1), bool)
                new_frontier[r][req_idx[r]] = ri
                packed[r][0][req_idx[r]] = np.where(keep, ri, default_node)
                packed[r][1][req_idx[r]] = np.where(keep, rw, 0)
                packed[r][2][req_idx[r]] = np.where(keep, rt, -1)
        for r in range(N):
            frontier[r] = new_frontier[r].reshape(-1)
            res[r][0].append(packed[r][0].reshape(-1)); res[r][1].append(packed[r][1].reshape(-1)); res[r][2].append(packed[r][2].reshape(-1))
    return res


def sage_mean_sharded(full_oracle, nbr_ids, rows, count, dim, N, me, P=None):
    """Association pinned for the fused sharded SAGE mean (eu_sym_sage_mean): owner o sums the feature rows of ITS ids per
    destination, j ascending, in f32; the requester adds the N partial rows in rank order and divides by
    f32(count) + 1e-7 (tf_euler/python/euler_ops/mp_ops.py:65-69).  full_oracle: OracleGraph of the WHOLE graph."""
    from euler_b200.sharded import owner_of
    ids = np.asarray(nbr_ids, dtype=np.int64).reshape(rows, count)
    feats = full_oracle.op_get_dense_feature(ids.reshape(-1), dim).reshape(rows, count, dim).astype(np.float32)
    own = owner_of(ids.reshape(-1), P or N, N, me).reshape(rows, count)
    total = np.zeros((rows, dim), np.float32)
    for o in range(N):
        part = np.zeros((rows, dim), np.float32)
        for j in range(count):
            part = (part + np.where((own[:, j] == o)[:, None], feats[:, j, :], np.float32(0))).astype(np.float32)
        total = (total + part).astype(np.float32)
    denom = np.float32(np.float32(count) + np.float32(1e-7))
    return (total / denom).astype(np.float32)


def type_weight_sums(shard):
    return np.bincount(shard["node_type"], weights=shard["node_w"].astype(np.float64), minlength=shard["n_node_types"])


def simulate_sample_node(shards, count, node_types, shard_seeds, client_seeds, repeat=1):
    """Single-process restatement of ShardedGraph.sample_node: per rank the ids it receives (shard order)."""
    from euler_b200.sharded import ClientRng, shard_weight_table, split_sample_count
    N = len(shards)
    table = shard_weight_table(np.stack([type_weight_sums(s) for s in shards]))
    ogs = []
    for s in shards:
        og = graphs.oracle_graph(s)
        og.build_node_sampler(np.arange(len(s["ids"]), dtype=np.int64), s["n_node_types"])
        ogs.append(og)
    rngs = [po.Rng(sd) for sd in shard_seeds]
    clients = [ClientRng(cs) for cs in client_seeds]
    out = None
    for _ in range(repeat):
        splits = [split_sample_count(count, node_types, table, clients[r]) for r in range(N)]
        got = [[None] * N for _ in range(N)]
        for s in range(N):
            for r in range(N):
                got[r][s] = ogs[s].sample_node(node_types, splits[r][s], rngs[s]) if splits[r][s] else np.zeros(0, np.uint64)
        out = [np.concatenate(got[r]).astype(np.int64) for r in range(N)]
    return out
