"""Generates the golden vectors under tests/golden/ from the REFERENCE ITSELF (oracle/_ref: the
unmodified reference sources + a seedable random.cc).  Runs only where /root/reference exists:

    oracle/tools/make_tiny_fixture.sh /tmp/euler      # reference converter -> .dat files
    python tests/golden/make_golden.py

Outputs (committed):
  tiny_euler/            the .dat/.meta files written by the reference converter
  tiny_csr.npz           that graph as loaded by the reference's Graph::Init, exported via the shim
  golden_ops.npz         outputs of the reference for seeded op calls on the tiny graph and on
                         seeded synthetic graphs (tests/graphs.py::random_graph)
The reference's own deterministic test vectors for this path (mp_ops_test.py:30-94,
neighbor_ops_test.py:46-57, compact_weighted_collection_test.cc:43-55) are written out in
tests/test_oracle_golden.py directly.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import pyoracle as po  # noqa: E402
import graphs  # noqa: E402

# (name, random_graph kwargs) -- regenerated identically by the tests
SYNTH = {
    "s1": dict(seed=11, n=300, T=1, avg_deg=5, feat_dim=8),
    "s3": dict(seed=12, n=200, T=3, avg_deg=4, n_node_types=2, zero_w_frac=0.1, id_stride=7, id_base=5),
    "s5": dict(seed=13, n=150, T=5, avg_deg=3, n_node_types=3, empty_frac=0.3, hub=200),
}
NB_CASES = [([0], 10), ([1], 4), ([0, 1], 5), ([], 3), ([1, 0], 40), ([7], 2)]


def seeds_for(g, rs, n):
    ids = g["ids"]
    s = ids[rs.randint(0, len(ids), size=n)].astype(np.int64)
    s[::7] = 10 ** 9 + 7  # absent id
    s[1::11] = 0
    return s


def main():
    out = {}
    # ---- tiny graph through the reference loader
    tiny = os.path.join(HERE, "tiny_euler")
    g = po.RefGraph.load(tiny, "node", "node")
    csr = g.export_csr()
    n = len(csr["ids"])
    dims = [2, 3]  # dense_f3, dense_f4 (euler.meta)
    feat = np.zeros((n, sum(dims)), np.float32)
    for r in range(n):
        ends, vals = csr["f32_ends"][r], csr["f32_vals"][r]
        off = 0
        for s, d in enumerate(dims):
            b = 0 if s == 0 else ends[s - 1]
            feat[r, off:off + min(d, ends[s] - b)] = vals[b:b + min(d, ends[s] - b)]
            off += d
    map_order = g.node_ids_in_map_order()
    np.savez(os.path.join(HERE, "tiny_csr.npz"), ids=csr["ids"], node_type=csr["node_type"],
             node_w=csr["node_w"], T=csr["T"], grp_ptr=csr["grp_ptr"], nbr=csr["nbr"],
             cum_w=csr["cum_w"], grp_cum=csr["grp_cum"], feat=feat, feat_slot_dims=np.asarray(dims, np.int32),
             n_node_types=2, map_order=map_order)
    tiny_seeds = np.asarray([1, 2, 3, 4, 5, 6, 7, 3, 1, 0, 6, 6], np.int64)
    out["tiny_seeds"] = tiny_seeds
    for ci, (et, cnt) in enumerate(NB_CASES):
        g.seed(100 + ci)
        ids, w, t = g.op_sample_neighbor(tiny_seeds, et, cnt, -1)
        out["tiny_nb%d_ids" % ci], out["tiny_nb%d_w" % ci], out["tiny_nb%d_t" % ci] = ids, w, t
        out["tiny_nb%d_draws" % ci] = g.draws()
    g.seed(200)
    ids, ws, ts = g.op_sample_fanout(tiny_seeds, [[0, 1], [0, 1]], [3, 4], -1)
    for l in range(2):
        out["tiny_fan%d_ids" % l], out["tiny_fan%d_w" % l], out["tiny_fan%d_t" % l] = ids[l], ws[l], ts[l]
    g.seed(300)
    out["tiny_walk_n2v"] = g.op_random_walk(tiny_seeds, np.asarray([[0, 1]] * 6, np.int32), 0.5, 2.0, -1)
    g.seed(301)
    out["tiny_walk_uni"] = g.op_random_walk(tiny_seeds, np.asarray([[0, 1]] * 6, np.int32), 1.0, 1.0, -1)
    for t in (0, 1):
        sid, sw, prob, alias = g.sampler_tables(t)
        out["tiny_sampler%d_ids" % t], out["tiny_sampler%d_prob" % t], out["tiny_sampler%d_alias" % t] = sid, prob, alias
    g.seed(400)
    out["tiny_sn_t0"] = g.sample_node([0], 64)
    g.seed(401)
    out["tiny_sn_all"] = g.sample_node([-1], 64)
    g.seed(402)
    out["tiny_sn_01"] = g.sample_node([0, 1], 64)
    f, lens = g.get_dense_feature(np.asarray([1, 9, 4], np.uint64), 1, 3)
    out["tiny_feat_f4"] = f

    # ---- seeded synthetic graphs through Node::Init
    for name, kw in SYNTH.items():
        sg = graphs.random_graph(**kw)
        rg = graphs.ref_graph(sg)
        rs = np.random.RandomState(kw["seed"] + 1000)
        seeds = seeds_for(sg, rs, 257)
        out[name + "_seeds"] = seeds
        T = sg["T"]
        cases = [([0], 10), (list(range(T)), 25), ([T - 1], 3)] + ([([0, T - 1], 7), ([1, 0], 33)] if T > 2 else [])
        out[name + "_ncases"] = len(cases)
        for ci, (et, cnt) in enumerate(cases):
            rg.seed(500 + ci)
            ids, w, t = rg.op_sample_neighbor(seeds, et, cnt, -1)
            out["%s_nb%d_et" % (name, ci)] = np.asarray(et, np.int32)
            out["%s_nb%d_cnt" % (name, ci)] = cnt
            out["%s_nb%d_ids" % (name, ci)], out["%s_nb%d_w" % (name, ci)], out["%s_nb%d_t" % (name, ci)] = ids, w, t
            out["%s_nb%d_draws" % (name, ci)] = rg.draws()
        rg.seed(600)
        ets = [[0]] * 2 if T == 1 else [[0, T - 1], [T - 1, 0]]
        ids, ws, ts = rg.op_sample_fanout(seeds, ets, [5, 3], -1)
        out[name + "_fan_et"] = np.asarray(ets, np.int32)
        for l in range(2):
            out["%s_fan%d_ids" % (name, l)], out["%s_fan%d_w" % (name, l)], out["%s_fan%d_t" % (name, l)] = ids[l], ws[l], ts[l]
        rg.seed(700)
        wet = np.asarray([list(range(T))] * 8, np.int32)
        out[name + "_walk_n2v"] = rg.op_random_walk(seeds[:64], wet, 0.5, 2.0, -1)
        rg.seed(701)
        out[name + "_walk_uni"] = rg.op_random_walk(seeds[:64], wet, 1.0, 1.0, -1)
        out[name + "_map_order"] = rg.node_ids_in_map_order()
        rg.seed(800)
        out[name + "_sn_all"] = rg.sample_node([-1], 500)
        rg.seed(801)
        out[name + "_sn_t0"] = rg.sample_node([0], 500)
    np.savez_compressed(os.path.join(HERE, "golden_ops.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
