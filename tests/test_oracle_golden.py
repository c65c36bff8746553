"""CPU: the C restatement (oracle/) against (a) the golden vectors produced by the reference itself
(tests/golden/make_golden.py) and (b) the deterministic vectors the reference's own tests hold for
this path.  This is what pins the oracle (parity status: pinned)."""
import numpy as np
import pytest

import cases
import graphs
from oracle import pyoracle as po


def test_rng_closed_form_matches_libstdcxx_known_answers():
    # minstd_rand0 known answer: 10000th value from seed 1 is 1043618065 (C++11 [rand.predef])
    r = po.Rng(1)
    x = 1
    for _ in range(10000):
        x = x * 16807 % 2147483647
    assert x == 1043618065
    for _ in range(5000):
        r.uniform()
    assert r.x == 1043618065 and r.draws == 5000
    # seed 0 -> state 1; seed m -> 1 (Appendix A-13)
    assert po.Rng(0).x == 1 and po.Rng(2147483647).x == 1 and po.Rng(2147483648).x == 1


def test_tiny_graph_export_matches_reference_tests():
    g = graphs.load_tiny_csr()
    # fixture adjacency quoted in SURVEY Appendix A-1 / graph_test.cc:68-76: node 1: nb=[2,4,3], cum=[2,6,9], gidx=[2,3]
    assert list(g["ids"]) == [1, 2, 3, 4, 5, 6]
    assert list(g["grp_ptr"][:3]) == [0, 2, 3]
    assert list(g["nbr"][:3]) == [2, 4, 3]
    assert list(g["cum_w"][:3]) == [2.0, 6.0, 9.0]
    og = graphs.oracle_graph(g)
    # neighbor_ops_test.py:46-57 get_full_neighbor of nodes [1,2], edge types [0,1]
    lens, ids, w, t = og.get_full_neighbor([1, 2], [0, 1])
    assert list(lens) == [3, 2]
    assert list(ids) == [2, 4, 3, 3, 5]
    assert list(w) == [2.0, 4.0, 3.0, 3.0, 5.0]
    assert list(t) == [0, 0, 1, 1, 1]


def test_mp_ops_reference_vectors():
    # tf_euler/python/euler_ops/mp_ops_test.py:30-94
    x = np.asarray([[1., 2.], [3., 4.], [5., 6.]], np.float32)
    idx = [1, 0, 1]
    assert po.scatter_add(x, idx, 2).tolist() == [[3., 4.], [6., 8.]]
    assert np.abs(po.scatter_mean(x, idx, 2) - [[3., 4.], [3., 4.]]).sum() < 1e-6
    x2 = np.asarray([[1., 6.], [3., 4.], [5., 2.]], np.float32)
    assert po.scatter_max(x2, idx, 2).tolist() == [[3., 4.], [5., 6.]]
    assert po.gather(x, [1, 0, 1, 2]).tolist() == [[3., 4.], [1., 2.], [3., 4.], [5., 6.]]
    # rows nobody scatters to: add -> 0, max -> -1e9 (scatter_op.cc:47,80)
    assert po.scatter_add(x, idx, 3)[2].tolist() == [0., 0.]
    assert po.scatter_max(x2, idx, 3)[2].tolist() == [np.float32(-1e9)] * 2


def test_cwc_get_weights_vector():
    # compact_weighted_collection_test.cc:43-55: ids 0..4 weights 1,2,3,4,5 -> Get(i) = (i, w_i)
    cum, _ = po.build_cum(np.asarray([0, 5], np.int64), np.asarray([1, 2, 3, 4, 5], np.float32), 1, 1)
    assert cum.tolist() == [1, 3, 6, 10, 15]


def test_zero_weight_entries_never_drawn():
    # compact_weighted_collection_test.cc:58-82 (zero-weight ids at both ends must never be sampled)
    ids = np.arange(6, dtype=np.int64)
    w = np.asarray([0, 1, 0, 2, 3, 0], np.float32)
    r = po.Rng(99)
    out_ids = np.zeros(200000, np.int64)
    out_w = np.zeros(200000, np.float32)
    po.lib().eo_cwc_sample(ids, w, 6, 200000, r.ref, out_ids, out_w)
    assert set(np.unique(out_ids)) == {1, 3, 4}
    frac = np.bincount(out_ids, minlength=6)[[1, 3, 4]] / 200000.0
    assert np.abs(frac - np.asarray([1, 2, 3]) / 6.0).max() < 0.01


def test_random_select_closed_form_equals_literal_search():
    """The CUDA kernels use min(end, first j with cum[j] > r); the reference does a 3-way binary
    search with fall-through.  Exhaustive check incl. zero-width intervals and r >= limit_end."""
    rs = np.random.RandomState(5)
    L = po.lib()
    for trial in range(300):
        n = rs.randint(1, 40)
        w = rs.randint(0, 4, size=n).astype(np.float32) * np.float32(0.7)
        if trial % 3 == 0:
            w[rs.rand(n) < 0.5] = 0
        cum = np.cumsum(w, dtype=np.float32)
        b = rs.randint(0, n)
        e = rs.randint(b, n)
        for s in range(20):
            r1, r2 = po.Rng(trial * 100 + s), po.Rng(trial * 100 + s)
            assert L.eo_random_select(cum, b, e, r1.ref) == L.eo_random_select_closed(cum, b, e, r2.ref)
    # forced r >= limit_end: inexact f32 subtraction makes diff too large
    cum = np.asarray([16777216.0, 16777218.0, 16777220.0], np.float32)
    for s in range(2000):
        r1, r2 = po.Rng(s), po.Rng(s)
        assert L.eo_random_select(cum, 1, 2, r1.ref) == L.eo_random_select_closed(cum, 1, 2, r2.ref)


def test_oracle_replays_reference_golden_tiny():
    g = graphs.load_tiny_csr()
    cases.replay_tiny(cases.OracleBackend(g, g["map_order"]))


@pytest.mark.parametrize("name", sorted(cases.SYNTH))
def test_oracle_replays_reference_golden_synth(name):
    g = graphs.random_graph(**cases.SYNTH[name])
    cases.replay_synth(name, cases.OracleBackend(g, cases.golden()[name + "_map_order"]))


def test_tiny_dense_feature_golden():
    g = graphs.load_tiny_csr()
    og = po.OracleGraph(g["ids"], g["node_type"], g["node_w"], g["T"], g["grp_ptr"], g["nbr"], g["cum_w"],
                        g["grp_cum"], np.ascontiguousarray(g["feat"][:, 2:5]))
    f = og.op_get_dense_feature([1, 9, 4], 3)
    cases.eq(f, cases.golden()["tiny_feat_f4"], "dense f4")
    # graph_test.cc / node_test.cc: node 1 f4 = [1.3,1.4,1.5]
    assert np.allclose(f[0], [1.3, 1.4, 1.5]) and not f[1].any()


def test_shard_routing():
    # euler/core/kernels/id_split_op.cc:46-49
    assert [po.shard_of(i, 8, 2) for i in range(10)] == [(i % 8) % 2 for i in range(10)]


def test_rmat_generator_restatement_invariants_and_pin():
    """oracle/rmat_gen.c (the host restatement of the device graph generator): structure, determinism across thread
    counts, and a pinned digest so that a silent change of either generator is caught on CPU (the GPU test compares the
    device generator against this one array by array)."""
    import hashlib
    a = po.rmat_graph(20_000, 150_000, seed=42, feat_dim=8, feat_seed=7, threads=1)
    b = po.rmat_graph(20_000, 150_000, seed=42, feat_dim=8, feat_seed=7, threads=5)
    for k in ("ids", "grp_ptr", "nbr", "cum_w", "feat"):
        assert np.array_equal(a[k], b[k]), k
    ptr, nbr, cum = a["grp_ptr"], a["nbr"], a["cum_w"]
    assert ptr[0] == 0 and ptr[-1] == 150_000 and np.all(np.diff(ptr) >= 0)
    assert nbr.min() >= 1 and nbr.max() <= 20_000
    rows = np.repeat(np.arange(20_000), np.diff(ptr))
    same = rows[1:] == rows[:-1]
    assert np.all(nbr[1:][same] >= nbr[:-1][same])                      # adjacency sorted by neighbor id
    w = np.diff(cum, prepend=np.float32(0)); w[ptr[:-1][np.diff(ptr) > 0]] = cum[ptr[:-1][np.diff(ptr) > 0]]
    assert w.min() > 0.99 and w.max() < 11.0                            # 1 + (h % 100) / 10, up to f32 rounding of the prefix
    assert np.abs(a["feat"]).max() <= 1.0
    h = hashlib.sha256()
    for k in ("grp_ptr", "nbr", "cum_w", "feat"):
        h.update(np.ascontiguousarray(a[k]).tobytes())
    assert h.hexdigest() == RMAT_DIGEST, h.hexdigest()


RMAT_DIGEST = "f0374b5e0a692cfbe72b1ece86e00b32787a7b1b4ff12d355e454958e95cc3b6"
