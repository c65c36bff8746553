"""Test helpers: synthetic graphs as plain numpy CSR arrays + builders for the oracle, the
reference shim and the CUDA product.  (Test infrastructure; imports oracle/.)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import pyoracle as po  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def random_graph(seed, n, T=1, avg_deg=6, n_node_types=1, zero_w_frac=0.0, empty_frac=0.1,
                 feat_dim=0, id_stride=1, id_base=1, hub=0, dup_edges=True, sorted_adj=True):
    """Random multigraph.  ids = id_base + id_stride * row (stride > 1 -> sparse ids -> hash lookups).
    hub > 0 adds one row with `hub` neighbours.  Weights are small multiples of 0.1 (f32)."""
    rng = np.random.RandomState(seed)
    ids = (id_base + id_stride * np.arange(n)).astype(np.uint64)
    deg = rng.poisson(avg_deg, size=(n, T))
    deg[rng.rand(n, T) < empty_frac] = 0
    if hub and n > 0:
        deg[rng.randint(n), rng.randint(T)] = hub
    grp_ptr = np.zeros(n * T + 1, np.int64)
    grp_ptr[1:] = np.cumsum(deg.reshape(-1))
    E = int(grp_ptr[-1])
    nbr = ids[rng.randint(0, n, size=E)] if n > 0 else np.zeros(0, np.uint64)
    if not dup_edges or sorted_adj:
        for k in range(n * T):
            b, e = grp_ptr[k], grp_ptr[k + 1]
            nbr[b:e] = np.sort(nbr[b:e])
    w = (1 + rng.randint(0, 100, size=E)).astype(np.float32) / np.float32(10)
    if zero_w_frac > 0:
        w[rng.rand(E) < zero_w_frac] = 0
    node_type = rng.randint(0, n_node_types, size=n).astype(np.int32)
    node_w = (1 + rng.randint(0, 50, size=n)).astype(np.float32) / np.float32(4)
    feat = rng.uniform(-1, 1, size=(n, feat_dim)).astype(np.float32) if feat_dim else None
    cum_w, grp_cum = po.build_cum(grp_ptr, w, n, T)
    return dict(ids=ids, node_type=node_type, node_w=node_w, T=T, grp_ptr=grp_ptr, nbr=nbr, w=w,
                cum_w=cum_w, grp_cum=grp_cum, feat=feat, n_node_types=n_node_types)


def oracle_graph(g):
    return po.OracleGraph(g["ids"], g["node_type"], g["node_w"], g["T"], g["grp_ptr"], g["nbr"],
                          g["cum_w"], g["grp_cum"], g.get("feat"))


def ref_graph(g, sampler=True):
    return po.RefGraph.build(g["ids"], g["node_type"], g["node_w"], g["T"], g["grp_ptr"], g["nbr"],
                             g["w"], g["n_node_types"], g.get("feat"), sampler)


def cuda_graph(g, raw_weights=False, sampler_order=None, device=0):
    import euler_b200
    kw = dict(w=g["w"]) if raw_weights else dict(cum_w=g["cum_w"], grp_cum=g["grp_cum"] if g["T"] > 1 else None)
    return euler_b200.Graph.from_csr(g["ids"], g["grp_ptr"], g["nbr"], n_edge_types=g["T"],
                                     node_type=g["node_type"], node_w=g["node_w"],
                                     n_node_types=g["n_node_types"], feat=g.get("feat"),
                                     sampler_order=sampler_order, device=device, **kw)


def load_tiny_csr():
    """tests/golden/tiny_csr.npz: the tools/test_data graph as loaded by the REFERENCE's own loader
    and exported through oracle/_ref (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "tiny_csr.npz"))
    g = {k: z[k] for k in z.files}
    g["T"] = int(g["T"])
    g["n_node_types"] = int(g["n_node_types"])
    return g
