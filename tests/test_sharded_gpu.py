"""GPU: sharding kernels on one GPU, and (when >= 2 GPUs are visible) the 2-rank NCCL run."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
import graphs

pytestmark = pytest.mark.gpu


def test_bucket_and_merge_kernels_vs_numpy():
    from euler_b200.sharded import CudaShardOps, owner_of
    g = graphs.random_graph(seed=3, n=100, T=1)
    ops = CudaShardOps(graphs.cuda_graph(g))
    rs = np.random.RandomState(0)
    for rows, P, N in [(0, 4, 4), (1, 8, 2), (1000, 8, 8), (70001, 12, 3), (25600, 64, 64)]:
        ids = rs.randint(0, 1 << 40, size=rows).astype(np.int64)
        ids[::5] = 0
        if rows > 3:
            ids[3] = -1
        me = N - 1
        s_ids, src, counts = ops.bucket(ops.to_dev(ids, torch.int64), P, N, me)
        own = owner_of(ids.astype(np.uint64), P, N, me).astype(np.int64)
        order = np.argsort(own, kind="stable")
        cases.eq(s_ids.cpu().numpy(), ids[order], "sorted ids rows=%d" % rows)
        cases.eq(src.cpu().numpy(), order.astype(np.int32), "src index")
        cases.eq(counts.cpu().numpy(), np.bincount(own, minlength=N).astype(np.int64), "counts")
        if rows == 0:
            continue
        c = 7
        r_ids = rs.randint(0, 50, size=(rows, c)).astype(np.int64)
        r_ids[::3, 0] = 0
        r_w = rs.rand(rows, c).astype(np.float32)
        r_t = rs.randint(0, 4, size=(rows, c)).astype(np.int32)
        packed = torch.empty(rows * c * 2, dtype=torch.int64, device="cuda")
        d_ids, d_w, d_t = ops.to_dev(r_ids.reshape(-1), torch.int64), ops.to_dev(r_w.reshape(-1), torch.float32), ops.to_dev(r_t.reshape(-1), torch.int32)
        ops.check(ops.lib.eu_shard_pack_sample(ops._stream(), d_ids.data_ptr(), d_w.data_ptr(), d_t.data_ptr(), rows * c, packed.data_ptr()))
        eng, o_ids, o_w, o_t = ops.merge_sample(packed, src, rows, c, -9)
        keep = r_ids[:, :1] != 0
        want = np.zeros((rows, c), np.int64); want[order] = r_ids
        cases.eq(eng.cpu().numpy().reshape(rows, c), want, "engine frontier")
        want[order] = np.where(keep, r_ids, -9)
        cases.eq(o_ids.cpu().numpy().reshape(rows, c), want, "packed ids")
        ww = np.zeros((rows, c), np.float32); ww[order] = np.where(keep, r_w, 0)
        cases.eq(o_w.cpu().numpy().reshape(rows, c), ww, "packed w")
        tt = np.zeros((rows, c), np.int32); tt[order] = np.where(keep, r_t, -1)
        cases.eq(o_t.cpu().numpy().reshape(rows, c), tt, "packed t")
        for D in (3, 64):
            x = rs.randn(rows, D).astype(np.float32)
            out = ops.merge_rows(ops.to_dev(x.reshape(-1), torch.float32), src, rows, D)
            wx = np.zeros_like(x); wx[order] = x
            cases.eq(out.cpu().numpy(), wx, "merge rows D=%d" % D)


def test_rmat_shards_union_is_the_full_graph():
    import euler_b200
    n, E, N = 50000, 400000, 4
    full = euler_b200.Graph.rmat(n, E, feat_dim=8).export()
    seen = 0
    for s in range(N):
        ex = euler_b200.Graph.rmat_shard(n, E, s, N, feat_dim=8).export()
        assert (ex["ids"] % N == s).all()
        rows = (ex["ids"] - 1).astype(np.int64)
        deg = np.diff(ex["grp_ptr"])
        assert np.array_equal(deg, np.diff(full["grp_ptr"])[rows])
        for r in list(range(0, len(rows), 97)) + [int(np.argmax(deg))]:
            b, e = ex["grp_ptr"][r], ex["grp_ptr"][r + 1]
            fb, fe = full["grp_ptr"][rows[r]], full["grp_ptr"][rows[r] + 1]
            assert np.array_equal(ex["nbr"][b:e], full["nbr"][fb:fe]) and np.array_equal(ex["cum_w"][b:e], full["cum_w"][fb:fe])
        cases.eq(ex["feat"], full["feat"][rows], "features of shard %d" % s)
        seen += len(rows)
    assert seen == n
    # lookups through the strided-dense id map
    sh = euler_b200.Graph.rmat_shard(n, E, 1, N, feat_dim=8)
    euler_b200.set_graph(sh)
    (f,) = euler_b200.get_dense_feature([1, 2, 5, n + 1, 0], [0], [8])
    f = f.cpu().numpy()
    assert np.array_equal(f[0], full["feat"][0]) and np.array_equal(f[2], full["feat"][4])
    assert not f[1].any() and not f[3].any() and not f[4].any()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_nccl_run_matches_oracle():
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(here, "run_sharded_gpu.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "SHARDED_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
