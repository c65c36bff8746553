"""TEST INFRASTRUCTURE -- ctypes bindings for the CPU oracle.

Two libraries, both test-only (see oracle/euler_oracle.h):
  * ``oracle/libeuler_oracle.so``      plain-C restatement (always buildable: gcc only)
  * ``oracle/_ref/libeuler_ref.so``    the unmodified reference sources + shim (built where
                                       /root/reference exists; travels to the GPU box prebuilt)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under euler_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libeuler_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libeuler_ref.so")

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C restatement (and, if /root/reference exists, the reference shim)."""
    srcs = [os.path.join(_HERE, f) for f in ("euler_oracle.c", "rmat_gen.c", "euler_oracle.h")]
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/euler/core/api/api.cc"):
        subprocess.check_call(["make", "-C", _HERE, "-j8", "ref"], stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(REF_SO)


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _ptr_array(arrs, ctype):
    t = (C.POINTER(ctype) * len(arrs))()
    for i, a in enumerate(arrs):
        t[i] = a.ctypes.data_as(C.POINTER(ctype))
    return t


# --------------------------------------------------------------------------- C restatement
_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(ORACLE_SO)
        L.eo_global_seed.argtypes = [C.c_uint64]
        L.eo_global_draws.restype = C.c_uint64
        L.eo_global_uniform.restype = C.c_double
        L.eo_global_state.restype = C.c_uint64
        L.eo_global_set_state.argtypes = [C.c_uint64, C.c_uint64]
        L.eo_graph_create.restype = C.c_void_p
        L.eo_graph_create.argtypes = [C.c_int64, C.c_int32, u64p, i32p, f32p, i64p, u64p, f32p, f32p,
                                      C.c_int32, C.c_void_p]
        L.eo_graph_destroy.argtypes = [C.c_void_p]
        L.eo_graph_row.restype = C.c_int64
        L.eo_graph_row.argtypes = [C.c_void_p, C.c_uint64]
        L.eo_build_cum.argtypes = [C.c_int64, C.c_int32, i64p, f32p, f32p, f32p]
        L.eo_random_select.restype = C.c_int64
        L.eo_random_select.argtypes = [f32p, C.c_int64, C.c_int64, C.c_void_p]
        L.eo_random_select_closed.restype = C.c_int64
        L.eo_random_select_closed.argtypes = [f32p, C.c_int64, C.c_int64, C.c_void_p]
        L.eo_seed.argtypes = [C.c_void_p, C.c_uint64]
        L.eo_uniform.restype = C.c_double
        L.eo_uniform.argtypes = [C.c_void_p]
        L.eo_cwc_sample.argtypes = [i64p, f32p, C.c_int64, C.c_int64, C.c_void_p, i64p, f32p]
        L.eo_alias_build.argtypes = [f32p, C.c_int64, f32p, i64p]
        L.eo_fwc_build.argtypes = [f32p, C.c_int64, f32p, i64p, C.POINTER(C.c_float)]
        L.eo_alias_next.restype = C.c_int64
        L.eo_alias_next.argtypes = [f32p, i64p, C.c_int64, C.c_void_p]
        L.eo_sample_neighbor.argtypes = [C.c_void_p, u64p, C.c_int64, i32p, C.c_int32, C.c_int32,
                                         C.c_void_p, u64p, f32p, i32p, i32p]
        L.eo_get_full_neighbor.restype = C.c_int64
        L.eo_get_full_neighbor.argtypes = [C.c_void_p, u64p, C.c_int64, i32p, C.c_int32, C.c_int64,
                                           i64p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.eo_node_sampler_create.restype = C.c_void_p
        L.eo_node_sampler_create.argtypes = [C.c_void_p, i64p, C.c_int64, C.c_int32]
        L.eo_node_sampler_destroy.argtypes = [C.c_void_p]
        L.eo_node_sampler_size.restype = C.c_int64
        L.eo_node_sampler_size.argtypes = [C.c_void_p, C.c_int32]
        L.eo_node_sampler_export.argtypes = [C.c_void_p, C.c_int32, u64p, f32p, f32p, i64p]
        L.eo_sample_node.restype = C.c_int64
        L.eo_sample_node.argtypes = [C.c_void_p, i32p, C.c_int32, C.c_int32, C.c_void_p, u64p]
        L.eo_op_sample_neighbor.argtypes = [C.c_void_p, i64p, C.c_int64, i32p, C.c_int32, C.c_int32,
                                            C.c_int64, i64p, f32p, i32p]
        L.eo_op_sample_fanout.argtypes = [C.c_void_p, i64p, C.c_int64, i32p, C.c_int32, i32p,
                                          C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.eo_op_random_walk.argtypes = [C.c_void_p, i64p, C.c_int64, i32p, C.c_int32, C.c_int32,
                                        C.c_float, C.c_float, C.c_int64, i64p]
        L.eo_op_get_dense_feature.argtypes = [C.c_void_p, i64p, C.c_int64, C.c_int32, f32p]
        L.eo_gather.argtypes = [f32p, C.c_int64, i32p, C.c_int64, f32p]
        for f in (L.eo_scatter_add, L.eo_scatter_max, L.eo_scatter_mean):
            f.argtypes = [f32p, C.c_int64, i32p, C.c_int64, C.c_int64, f32p]
        L.eo_shard_of.restype = C.c_int32
        L.eo_shard_of.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
        L.eo_bench_fanout.restype = C.c_double
        L.eo_bench_fanout.argtypes = [C.c_void_p, i64p, C.c_int64, C.c_int64, i32p, C.c_int32, i32p,
                                      C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        L.eo_bench_step.restype = C.c_double
        L.eo_bench_step.argtypes = [C.c_void_p, i64p, C.c_int64, C.c_int64, i32p, C.c_int32, i32p,
                                    C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        L.eo_rmat_csr.restype = C.c_int
        L.eo_rmat_csr.argtypes = [C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_int32,
                                  C.c_int32, C.c_int32, u64p, i32p, f32p, i64p, u64p, f32p, C.c_void_p]
        L.eo_rmat_feat_rows.argtypes = [i64p, C.c_int64, C.c_int64, C.c_int32, C.c_uint64, f32p]
        L.eo_rmat_feat_full.argtypes = [C.c_int64, C.c_int32, C.c_uint64, C.c_int32, f32p]
        _lib = L
    return _lib


def rmat_graph(n_nodes, n_edges, a=0.57, b=0.19, c=0.19, seed=42, feat_dim=0, feat_seed=7, T=1, NT=1, threads=None):
    """Host copy of euler_b200.Graph.rmat / rmat_hetero (oracle/rmat_gen.c): the same dict Graph.export() returns."""
    threads = threads or max(1, len(os.sched_getaffinity(0)))
    out = dict(ids=np.zeros(n_nodes, np.uint64), node_type=np.zeros(n_nodes, np.int32), node_w=np.zeros(n_nodes, np.float32),
               grp_ptr=np.zeros(n_nodes * T + 1, np.int64), nbr=np.zeros(n_edges, np.uint64),
               cum_w=np.zeros(n_edges, np.float32), grp_cum=np.zeros(n_nodes * T, np.float32) if T > 1 else None, feat=None, T=T)
    rc = lib().eo_rmat_csr(n_nodes, n_edges, a, b, c, seed, T, NT, threads, out["ids"], out["node_type"], out["node_w"],
                           out["grp_ptr"], out["nbr"], out["cum_w"], None if T == 1 else out["grp_cum"].ctypes.data)
    if rc:
        raise RuntimeError("eo_rmat_csr failed")
    if feat_dim:
        out["feat"] = np.zeros((n_nodes, feat_dim), np.float32)
        lib().eo_rmat_feat_full(n_nodes, feat_dim, feat_seed, threads, out["feat"].reshape(-1))
    return out


def rmat_feat_rows(ids, n_nodes, dim, feat_seed=7):
    """feature rows of the synthetic graph for `ids` (zeros for ids that are not nodes), f32[len(ids), dim]"""
    ids = _arr(ids, np.int64).reshape(-1)
    out = np.zeros((len(ids), dim), np.float32)
    lib().eo_rmat_feat_rows(ids, len(ids), n_nodes, dim, feat_seed, out.reshape(-1))
    return out


class Rng(C.Structure):
    _fields_ = [("x", C.c_uint64), ("draws", C.c_uint64)]

    def __init__(self, seed=1):
        super().__init__()
        lib().eo_seed(C.byref(self), seed)

    def uniform(self):
        return lib().eo_uniform(C.byref(self))

    @property
    def ref(self):
        return C.byref(self)


def seed(s):
    lib().eo_global_seed(s)


def draws():
    return lib().eo_global_draws()


def get_state():
    return lib().eo_global_state(), lib().eo_global_draws()


def set_state(st):
    lib().eo_global_set_state(st[0], st[1])


def build_cum(grp_ptr, w, n, T):
    """Node::Init accumulation: raw weights -> (cum_w, grp_cum)."""
    w = _arr(w, np.float32)
    cum = np.zeros_like(w)
    gc = np.zeros(n * T, np.float32)
    lib().eo_build_cum(n, T, _arr(grp_ptr, np.int64), w, cum, gc)
    return cum, gc


class OracleGraph:
    """CSR graph for the C restatement.  All arrays are kept alive on the object."""

    def __init__(self, ids, node_type, node_w, T, grp_ptr, nbr, cum_w, grp_cum, feat=None):
        self.n = len(ids)
        self.T = int(T)
        self.ids = _arr(ids, np.uint64)
        self.node_type = _arr(node_type, np.int32)
        self.node_w = _arr(node_w, np.float32)
        self.grp_ptr = _arr(grp_ptr, np.int64)
        self.nbr = _arr(nbr, np.uint64)
        self.cum_w = _arr(cum_w, np.float32)
        self.grp_cum = _arr(grp_cum, np.float32)
        assert self.grp_ptr.shape[0] == self.n * self.T + 1
        self.feat = None if feat is None else _arr(feat, np.float32)
        self.feat_dim = 0 if feat is None else self.feat.shape[1]
        self.h = lib().eo_graph_create(self.n, self.T, self.ids, self.node_type, self.node_w,
                                       self.grp_ptr, self.nbr, self.cum_w, self.grp_cum,
                                       self.feat_dim,
                                       None if feat is None else self.feat.ctypes.data)
        self._sampler = None

    def __del__(self):
        try:
            if self._sampler:
                lib().eo_node_sampler_destroy(self._sampler)
            lib().eo_graph_destroy(self.h)
        except Exception:
            pass

    # ---- api.cc level
    def sample_neighbor_api(self, ids, etypes, count, rng):
        ids = _arr(ids, np.uint64)
        et = _arr(etypes, np.int32)
        n = len(ids)
        o_ids = np.zeros((n, count), np.uint64)
        o_w = np.zeros((n, count), np.float32)
        o_t = np.zeros((n, count), np.int32)
        o_len = np.zeros(n, np.int32)
        lib().eo_sample_neighbor(self.h, ids, n, et, len(et), count, rng.ref, o_ids, o_w, o_t, o_len)
        return o_ids, o_w, o_t, o_len

    def get_full_neighbor(self, ids, etypes):
        ids = _arr(ids, np.uint64)
        et = _arr(etypes, np.int32)
        n = len(ids)
        lens = np.zeros(n, np.int64)
        tot = lib().eo_get_full_neighbor(self.h, ids, n, et, len(et), 0, lens, None, None, None)
        o_ids = np.zeros(max(tot, 1), np.uint64)
        o_w = np.zeros(max(tot, 1), np.float32)
        o_t = np.zeros(max(tot, 1), np.int32)
        lib().eo_get_full_neighbor(self.h, ids, n, et, len(et), tot, lens, o_ids.ctypes.data,
                                   o_w.ctypes.data, o_t.ctypes.data)
        return lens, o_ids[:tot], o_w[:tot], o_t[:tot]

    # ---- global node sampler
    def build_node_sampler(self, order_rows, n_types):
        order = _arr(order_rows, np.int64)
        self._sampler = lib().eo_node_sampler_create(self.h, order, len(order), n_types)
        self.n_node_types = n_types

    def node_sampler_tables(self, t):
        m = lib().eo_node_sampler_size(self._sampler, t)
        ids = np.zeros(m, np.uint64)
        w = np.zeros(m, np.float32)
        prob = np.zeros(m, np.float32)
        alias = np.zeros(m, np.int64)
        lib().eo_node_sampler_export(self._sampler, t, ids, w, prob, alias)
        return ids, w, prob, alias

    def sample_node(self, types, count, rng):
        types = _arr(np.atleast_1d(types), np.int32)
        out = np.zeros(max(count, 1), np.uint64)
        m = lib().eo_sample_node(self._sampler, types, len(types), count, rng.ref, out)
        return out[:m]

    # ---- tf_euler op level (global stream: call oracle.seed() first)
    def op_sample_neighbor(self, nodes, etypes, count, default_node=-1):
        nodes = _arr(nodes, np.int64)
        et = _arr(etypes, np.int32)
        n = len(nodes)
        o_ids = np.zeros((n, count), np.int64)
        o_w = np.zeros((n, count), np.float32)
        o_t = np.zeros((n, count), np.int32)
        lib().eo_op_sample_neighbor(self.h, nodes, n, et, len(et), count, default_node, o_ids, o_w, o_t)
        return o_ids, o_w, o_t

    def op_sample_fanout(self, nodes, etypes, counts, default_node=-1):
        """etypes: [L,K] array.  Returns lists of flat arrays per hop (ids, w, t)."""
        nodes = _arr(nodes, np.int64)
        et = _arr(etypes, np.int32).reshape(len(counts), -1)
        cs = _arr(counts, np.int32)
        n = len(nodes)
        ids, ws, ts = [], [], []
        rows = n
        for c in counts:
            rows *= c
            ids.append(np.zeros(rows, np.int64))
            ws.append(np.zeros(rows, np.float32))
            ts.append(np.zeros(rows, np.int32))
        lib().eo_op_sample_fanout(self.h, nodes, n, et, et.shape[1], cs, len(counts), default_node,
                                  _ptr_array(ids, C.c_int64), _ptr_array(ws, C.c_float),
                                  _ptr_array(ts, C.c_int32))
        return ids, ws, ts

    def op_random_walk(self, nodes, etypes, p, q, default_node=-1):
        """etypes: [L,K]."""
        nodes = _arr(nodes, np.int64)
        et = _arr(etypes, np.int32)
        L, K = et.shape
        out = np.zeros((len(nodes), L + 1), np.int64)
        lib().eo_op_random_walk(self.h, nodes, len(nodes), et, K, L, p, q, default_node, out)
        return out

    def op_get_dense_feature(self, nodes, dim):
        nodes = _arr(nodes, np.int64)
        out = np.zeros((len(nodes), dim), np.float32)
        lib().eo_op_get_dense_feature(self.h, nodes, len(nodes), dim, out)
        return out

    def bench_fanout(self, seeds, etypes, counts, n_threads, iters):
        seeds = _arr(seeds, np.int64)
        nb, B = seeds.shape
        et = _arr(etypes, np.int32).reshape(len(counts), -1)
        edges = C.c_int64(0)
        sec = lib().eo_bench_fanout(self.h, seeds, nb, B, et, et.shape[1], _arr(counts, np.int32),
                                    len(counts), n_threads, iters, C.byref(edges))
        return sec, edges.value


def _bench_args(seeds, etypes, counts):
    seeds = _arr(seeds, np.int64)
    et = _arr(etypes, np.int32).reshape(len(counts), -1)
    return seeds, et, _arr(counts, np.int32)


def oracle_bench_step(og, seeds, etypes, counts, dim, n_threads, iters):
    seeds, et, cs = _bench_args(seeds, etypes, counts)
    edges = C.c_int64(0)
    sec = lib().eo_bench_step(og.h, seeds, seeds.shape[0], seeds.shape[1], et, et.shape[1], cs, len(cs),
                              dim, n_threads, iters, C.byref(edges))
    return sec, edges.value


def ref_bench_step(seeds, etypes, counts, dim, n_threads, iters):
    seeds, et, cs = _bench_args(seeds, etypes, counts)
    edges = C.c_int64(0)
    sec = ref().ref_bench_step(seeds, seeds.shape[0], seeds.shape[1], et, et.shape[1], cs, len(cs), dim,
                               n_threads, iters, C.byref(edges))
    return sec, edges.value


def gather(params, idx):
    params = _arr(params, np.float32)
    idx = _arr(idx, np.int32)
    out = np.zeros((len(idx), params.shape[1]), np.float32)
    lib().eo_gather(params, params.shape[1], idx, len(idx), out)
    return out


def _scatter(fn, upd, idx, size):
    upd = _arr(upd, np.float32)
    idx = _arr(idx, np.int32)
    out = np.zeros((size, upd.shape[1]), np.float32)
    fn(upd, upd.shape[1], idx, len(idx), size, out)
    return out


def scatter_add(upd, idx, size):
    return _scatter(lib().eo_scatter_add, upd, idx, size)


def scatter_max(upd, idx, size):
    return _scatter(lib().eo_scatter_max, upd, idx, size)


def scatter_mean(upd, idx, size):
    return _scatter(lib().eo_scatter_mean, upd, idx, size)


def shard_of(i, num_partitions, shard_num):
    return lib().eo_shard_of(int(i), num_partitions, shard_num)


# --------------------------------------------------------------------------- reference shim
_ref = None


def ref():
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        R = C.CDLL(REF_SO)
        R.ref_seed.argtypes = [C.c_uint64]
        R.ref_draws.restype = C.c_uint64
        R.ref_uniform.restype = C.c_double
        R.ref_graph_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p]
        R.ref_graph_build.argtypes = [C.c_int64, u64p, i32p, f32p, C.c_int32, i64p, u64p, f32p,
                                      C.c_int32, C.c_int32, C.c_void_p, C.c_int32]
        R.ref_node_count.restype = C.c_int64
        R.ref_export_node_ids.argtypes = [u64p]
        R.ref_node_info.argtypes = [C.c_uint64] + [C.POINTER(C.c_int32), C.POINTER(C.c_float)] + \
            [C.POINTER(C.c_int32)] * 4
        R.ref_node_adj.argtypes = [C.c_uint64, i32p, u64p, f32p, f32p]
        R.ref_node_f32feat.argtypes = [C.c_uint64, i32p, f32p]
        R.ref_sampler_size.restype = C.c_int64
        R.ref_sampler_size.argtypes = [C.c_int32]
        R.ref_sampler_export.argtypes = [C.c_int32, u64p, f32p, f32p, i64p]
        R.ref_type_sampler_export.argtypes = [f32p, f32p, i64p]
        R.ref_random_select.restype = C.c_int64
        R.ref_random_select.argtypes = [f32p, C.c_int64, C.c_int64, C.c_int64]
        R.ref_cwc_sample.argtypes = [i64p, f32p, C.c_int64, C.c_int64, i64p, f32p]
        R.ref_alias_build.argtypes = [f32p, C.c_int64, f32p, i64p]
        R.ref_fwc_sample.argtypes = [u64p, f32p, C.c_int64, C.c_int64, u64p]
        R.ref_sample_neighbor.argtypes = [u64p, C.c_int64, i32p, C.c_int32, C.c_int32, u64p, f32p,
                                          i32p, i32p]
        R.ref_get_full_neighbor.restype = C.c_int64
        R.ref_get_full_neighbor.argtypes = [u64p, C.c_int64, i32p, C.c_int32, C.c_int64, i64p,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_sample_node.restype = C.c_int64
        R.ref_sample_node.argtypes = [i32p, C.c_int32, C.c_int32, u64p]
        R.ref_sample_edge.restype = C.c_int64
        R.ref_sample_edge.argtypes = [i32p, C.c_int32, C.c_int32, u64p]
        R.ref_get_dense_feature.argtypes = [u64p, C.c_int64, C.c_int32, C.c_int32, f32p, i32p]
        R.ref_op_sample_neighbor.argtypes = [i64p, C.c_int64, i32p, C.c_int32, C.c_int32, C.c_int64,
                                             i64p, f32p, i32p]
        R.ref_op_sample_fanout.argtypes = [i64p, C.c_int64, i32p, C.c_int32, i32p, C.c_int32,
                                           C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_op_random_walk.argtypes = [i64p, C.c_int64, i32p, C.c_int32, C.c_int32, C.c_float,
                                         C.c_float, C.c_int64, i64p]
        R.ref_bench_fanout.restype = C.c_double
        R.ref_bench_fanout.argtypes = [i64p, C.c_int64, C.c_int64, i32p, C.c_int32, i32p, C.c_int32,
                                       C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        R.ref_bench_step.restype = C.c_double
        R.ref_bench_step.argtypes = [i64p, C.c_int64, C.c_int64, i32p, C.c_int32, i32p, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
        R.ref_bench_feature.restype = C.c_double
        R.ref_bench_feature.argtypes = [i64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32]
        _ref = R
    return _ref


class RefGraph:
    """The reference's singleton euler::Graph behind the shim (one live instance per process)."""

    @staticmethod
    def load(directory, sampler_type="node", data_type="node"):
        rc = ref().ref_graph_load(directory.encode(), sampler_type.encode(), data_type.encode())
        if rc != 0:
            raise RuntimeError("reference Graph::Init failed for %s" % directory)
        return RefGraph()

    @staticmethod
    def build(ids, node_type, node_w, T, grp_ptr, nbr, w, n_node_types, feat=None, sampler=True):
        feat_c = None if feat is None else _arr(feat, np.float32)
        rc = ref().ref_graph_build(len(ids), _arr(ids, np.uint64), _arr(node_type, np.int32),
                                   _arr(node_w, np.float32), T, _arr(grp_ptr, np.int64),
                                   _arr(nbr, np.uint64), _arr(w, np.float32), n_node_types,
                                   0 if feat is None else feat_c.shape[1],
                                   None if feat is None else feat_c.ctypes.data, int(sampler))
        if rc != 0:
            raise RuntimeError("reference Node::Init failed")
        return RefGraph()

    def seed(self, s):
        ref().ref_seed(s)

    def draws(self):
        return ref().ref_draws()

    def node_ids_in_map_order(self):
        ids = np.zeros(ref().ref_node_count(), np.uint64)
        ref().ref_export_node_ids(ids)
        return ids

    def export_csr(self, ids=None):
        """Export the loaded graph as CSR arrays (rows in ascending id order unless ids given)."""
        R = ref()
        if ids is None:
            ids = np.sort(self.node_ids_in_map_order())
        n = len(ids)
        T = R.ref_edge_type_num()
        types = np.zeros(n, np.int32)
        nw = np.zeros(n, np.float32)
        grp_ptr = np.zeros(n * T + 1, np.int64)
        nbrs, cums, gcums, feats, fends = [], [], [], [], []
        t_, w_, g_, d_, s_, v_ = (C.c_int32(), C.c_float(), C.c_int32(), C.c_int32(), C.c_int32(),
                                  C.c_int32())
        for r, i in enumerate(ids):
            assert R.ref_node_info(int(i), C.byref(t_), C.byref(w_), C.byref(g_), C.byref(d_),
                                   C.byref(s_), C.byref(v_)) == 0
            types[r], nw[r] = t_.value, w_.value
            assert g_.value == T, "node %d carries %d edge groups, graph has %d" % (i, g_.value, T)
            ge = np.zeros(max(T, 1), np.int32)
            nb = np.zeros(max(d_.value, 1), np.uint64)
            cw = np.zeros(max(d_.value, 1), np.float32)
            gc = np.zeros(max(T, 1), np.float32)
            R.ref_node_adj(int(i), ge, nb, cw, gc)
            grp_ptr[r * T + 1:r * T + T + 1] = grp_ptr[r * T] + ge[:T]
            nbrs.append(nb[:d_.value]); cums.append(cw[:d_.value]); gcums.append(gc[:T])
            se = np.zeros(max(s_.value, 1), np.int32)
            fv = np.zeros(max(v_.value, 1), np.float32)
            R.ref_node_f32feat(int(i), se, fv)
            fends.append(se[:s_.value]); feats.append(fv[:v_.value])
        return dict(ids=np.asarray(ids, np.uint64), node_type=types, node_w=nw, T=T, grp_ptr=grp_ptr,
                    nbr=np.concatenate(nbrs) if nbrs else np.zeros(0, np.uint64),
                    cum_w=np.concatenate(cums) if cums else np.zeros(0, np.float32),
                    grp_cum=np.concatenate(gcums) if gcums else np.zeros(0, np.float32),
                    f32_ends=fends, f32_vals=feats)

    def sampler_tables(self, t):
        R = ref()
        m = R.ref_sampler_size(t)
        ids = np.zeros(m, np.uint64); w = np.zeros(m, np.float32)
        prob = np.zeros(m, np.float32); alias = np.zeros(m, np.int64)
        R.ref_sampler_export(t, ids, w, prob, alias)
        return ids, w, prob, alias

    def sample_neighbor_api(self, ids, etypes, count):
        ids = _arr(ids, np.uint64); et = _arr(etypes, np.int32)
        n = len(ids)
        o_ids = np.zeros((n, count), np.uint64); o_w = np.zeros((n, count), np.float32)
        o_t = np.zeros((n, count), np.int32); o_len = np.zeros(n, np.int32)
        ref().ref_sample_neighbor(ids, n, et, len(et), count, o_ids, o_w, o_t, o_len)
        return o_ids, o_w, o_t, o_len

    def get_full_neighbor(self, ids, etypes):
        ids = _arr(ids, np.uint64); et = _arr(etypes, np.int32)
        n = len(ids)
        lens = np.zeros(n, np.int64)
        tot = ref().ref_get_full_neighbor(ids, n, et, len(et), 0, lens, None, None, None)
        o_ids = np.zeros(max(tot, 1), np.uint64); o_w = np.zeros(max(tot, 1), np.float32)
        o_t = np.zeros(max(tot, 1), np.int32)
        ref().ref_get_full_neighbor(ids, n, et, len(et), tot, lens, o_ids.ctypes.data,
                                    o_w.ctypes.data, o_t.ctypes.data)
        return lens, o_ids[:tot], o_w[:tot], o_t[:tot]

    def sample_node(self, types, count):
        types = _arr(np.atleast_1d(types), np.int32)
        out = np.zeros(max(count, 1), np.uint64)
        m = ref().ref_sample_node(types, len(types), count, out)
        return out[:m]

    def sample_edge(self, types, count):
        """euler::SampleEdge: (n, 3) int64 array of (src, dst, type); n == 0 for several types (upstream quirk)"""
        types = _arr(np.atleast_1d(types), np.int32)
        out = np.zeros(max(count, 1) * 3, np.uint64)
        m = ref().ref_sample_edge(types, len(types), count, out)
        return out[:m * 3].astype(np.int64).reshape(m, 3)

    def get_dense_feature(self, ids, fid, dim):
        ids = _arr(ids, np.uint64)
        out = np.zeros((len(ids), dim), np.float32); lens = np.zeros(len(ids), np.int32)
        ref().ref_get_dense_feature(ids, len(ids), fid, dim, out, lens)
        return out, lens

    def op_sample_neighbor(self, nodes, etypes, count, default_node=-1):
        nodes = _arr(nodes, np.int64); et = _arr(etypes, np.int32)
        n = len(nodes)
        o_ids = np.zeros((n, count), np.int64); o_w = np.zeros((n, count), np.float32)
        o_t = np.zeros((n, count), np.int32)
        ref().ref_op_sample_neighbor(nodes, n, et, len(et), count, default_node, o_ids, o_w, o_t)
        return o_ids, o_w, o_t

    def op_sample_fanout(self, nodes, etypes, counts, default_node=-1):
        nodes = _arr(nodes, np.int64)
        et = _arr(etypes, np.int32).reshape(len(counts), -1)
        cs = _arr(counts, np.int32)
        ids, ws, ts = [], [], []
        rows = len(nodes)
        for c in counts:
            rows *= c
            ids.append(np.zeros(rows, np.int64)); ws.append(np.zeros(rows, np.float32))
            ts.append(np.zeros(rows, np.int32))
        ref().ref_op_sample_fanout(nodes, len(nodes), et, et.shape[1], cs, len(counts), default_node,
                                   _ptr_array(ids, C.c_int64), _ptr_array(ws, C.c_float),
                                   _ptr_array(ts, C.c_int32))
        return ids, ws, ts

    def op_random_walk(self, nodes, etypes, p, q, default_node=-1):
        nodes = _arr(nodes, np.int64); et = _arr(etypes, np.int32)
        L, K = et.shape
        out = np.zeros((len(nodes), L + 1), np.int64)
        ref().ref_op_random_walk(nodes, len(nodes), et, K, L, p, q, default_node, out)
        return out

    def bench_fanout(self, seeds, etypes, counts, n_threads, iters, with_dedup=True):
        seeds = _arr(seeds, np.int64)
        nb, B = seeds.shape
        et = _arr(etypes, np.int32).reshape(len(counts), -1)
        edges = C.c_int64(0)
        sec = ref().ref_bench_fanout(seeds, nb, B, et, et.shape[1], _arr(counts, np.int32),
                                     len(counts), n_threads, iters, int(with_dedup), C.byref(edges))
        return sec, edges.value

    def bench_feature(self, ids, dim, n_threads, iters):
        ids = _arr(ids, np.int64)
        return ref().ref_bench_feature(ids, len(ids), dim, n_threads, iters)
