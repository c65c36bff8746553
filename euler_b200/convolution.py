"""Message-passing blocks of tf_euler's convolutions over the euler_b200 mp ops (gather / scatter_*): the access patterns of
GCNConv and RelationConv.  The dense layers (tf.layers.Dense, the per-relation matmul) stay
with the framework (torch here): they are GEMMs, not part of the sampling / aggregation path.

  gcn_aggregate      tf_euler/python/convolution/gcn_conv.py:32-48   deg^-1/2 norms via scatter_add(ones) on both sides of
                                                                      edge_index, gather of x_j and of both norms, scatter_add
  relation_aggregate tf_euler/python/convolution/relation_conv.py:53-70  x_j gathered, transformed by its relation's matrix
                                                                      (unique -> gather -> matmul), scatter_mean to the targets
  sage_aggregate     tf_euler/python/convolution/sage_conv.py:33-38   gather(x, edge_index[1]) -> scatter_mean
"""
import torch

from . import ops


def gcn_norm(edge_index, size):
    """GCNConv.norm (gcn_conv.py:32-40): (deg_0 ** -0.5, deg_1 ** -0.5), deg_i = scatter_add(ones, edge_index[i], size[i])"""
    e = edge_index.shape[1]
    ones = torch.ones((e, 1), dtype=torch.float32, device=edge_index.device)
    return tuple(ops.scatter_add(ones, edge_index[i].to(torch.int32), int(size[i])) ** -0.5 for i in (0, 1))


def gcn_aggregate(x, edge_index, size):
    """GCNConv.__call__ up to (not including) the Dense layer (gcn_conv.py:42-55): x = (x_target, x_source) or one tensor for
    both sides; returns scatter_add(norm_i * norm_j * x_j, edge_index[0], size[0])."""
    x0, x1 = (x, x) if torch.is_tensor(x) else (x[0], x[1] if x[1] is not None else x[0])
    del x0
    idx0, idx1 = edge_index[0].to(torch.int32), edge_index[1].to(torch.int32)
    n0, n1 = gcn_norm(edge_index, size)
    x_j = ops.gather(x1, idx1)
    out = ops.gather(n0, idx0) * ops.gather(n1, idx1) * x_j
    return ops.scatter_add(out, idx0, int(size[0]))


def relation_aggregate(x, edge_index, size, edge_attr, matrix):
    """RelationConv.__call__ up to apply_node (relation_conv.py:53-70): matrix f32[num_relations, dim, fea_dim];
    out[i] = mean over edges e with target i of matrix[edge_attr[e]] @ x_source[edge_index[1][e]]."""
    x1 = x if torch.is_tensor(x) else (x[1] if x[1] is not None else x[0])
    idx0, idx1 = edge_index[0].to(torch.int32), edge_index[1].to(torch.int32)
    x_j = ops.gather(x1, idx1)
    rel, inv = torch.unique(edge_attr, return_inverse=True)          # tf.unique + the two gathers of apply_edge
    m = matrix[rel][inv]                                             # [E, dim, fea_dim]
    out = torch.matmul(m, x_j.unsqueeze(-1)).squeeze(-1)
    return ops.scatter_mean(out, idx0, int(size[0]))


def sage_aggregate(x, edge_index, size):
    """SAGEConv's neighbor mean (sage_conv.py:33-38): scatter_mean(gather(x_source, edge_index[1]), edge_index[0], size[0])"""
    x1 = x if torch.is_tensor(x) else (x[1] if x[1] is not None else x[0])
    return ops.scatter_mean(ops.gather(x1, edge_index[1].to(torch.int32)), edge_index[0].to(torch.int32), int(size[0]))
