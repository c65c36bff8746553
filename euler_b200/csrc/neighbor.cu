// Full-neighbor listing on the HBM-resident CSR.
// Reference semantics reproduced (file:line relative to /root/reference):
//   Node::GetFullNeighbor        euler/core/graph/node.cc:176-198  -- per requested edge type, in the order given
//                                (repeats repeat), every edge of that group as (id, weight, type)
//   euler::GetFullNeighbor       euler/core/api/api.cc:208-221     -- per node, missing node -> empty list
//   tf_euler GetFullNeighbor     tf_euler/kernels/get_full_neighbor_op.cc -- CSR-style (ragged) result
// Weights: the reference stores node-global cumulative weights and returns cum[j] - cum[j-1] (f32); same here.
//
// Three kernels, no host sync: per-node lengths -> cub inclusive scan -> one warp per node copies its groups.
#include <cub/device/device_scan.cuh>

#include "internal.h"

namespace eu {

struct ETList {
  int32_t K;
  int32_t v[EU_MAX_ETYPES];
};

__global__ void k_full_len(DevGraph g, const unsigned long long* __restrict__ nodes, int64_t B, ETList et,
                           long long* __restrict__ out_ptr /* [B+1]; [0] = 0, [i+1] = len(i) */) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) out_ptr[0] = 0;
  if (i >= B) return;
  const int64_t row = lookup_row(g, nodes[i]);
  long long len = 0;
  if (row >= 0) {
    const int64_t* gp = g.grp_ptr + row * g.T;
    for (int32_t k = 0; k < et.K; ++k) {
      const int32_t t = et.v[k];
      if (t >= 0 && t < g.T) len += gp[t + 1] - gp[t];
    }
  }
  out_ptr[i + 1] = len;
}

__global__ void __launch_bounds__(256) k_full_fill(DevGraph g, const unsigned long long* __restrict__ nodes, int64_t B, ETList et,
                                                   const long long* __restrict__ out_ptr, int64_t cap,
                                                   unsigned long long* __restrict__ out_ids, float* __restrict__ out_w,
                                                   int32_t* __restrict__ out_t) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < B; i += nwarps) {
    const int64_t row = lookup_row(g, nodes[i]);
    if (row < 0) continue;
    const int64_t* gp = g.grp_ptr + row * g.T;
    const int64_t base = gp[0];
    int64_t o = out_ptr[i];
    for (int32_t k = 0; k < et.K; ++k) {
      const int32_t t = et.v[k];
      if (t < 0 || t >= g.T) continue;
      const int64_t b = gp[t], e = gp[t + 1];
      for (int64_t j = b + lane; j < e; j += 32) {
        const int64_t pos = o + (j - b);
        if (pos < cap) {
          out_ids[pos] = g.nbr[j];
          out_w[pos] = __fsub_rn(g.cum_w[j], j == base ? 0.f : g.cum_w[j - 1]);
          out_t[pos] = t;
        }
      }
      o += e - b;
    }
  }
}

// euler::GetNodeType (api.cc:50-61): the node's type, DEFAULT_INT32 (= INT32_MIN, data_types.cc:23) when the node is absent
__global__ void k_node_type(DevGraph g, const unsigned long long* __restrict__ nodes, int64_t B, int32_t* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  const int64_t row = lookup_row(g, nodes[i]);
  out[i] = row >= 0 ? g.node_type[row] : (int32_t)0x80000000;
}

}  // namespace eu

using namespace eu;

extern "C" int eu_get_node_type(eu_ctx* c, const int64_t* nodes, int64_t B, int32_t* out) {
  if (!c || B < 0 || (B > 0 && (!nodes || !out))) { set_error("eu_get_node_type: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (B == 0) return EU_OK;
  k_node_type<<<(unsigned)ceil_div(B, 256), 256, 0, c->stream>>>(c->g->d, (const unsigned long long*)nodes, B, out);
  EU_LAUNCHED();
  return EU_OK;
}

extern "C" int eu_get_node_type_host(eu_ctx* c, const int64_t* nodes, int64_t B, int32_t* out) {
  if (!c || B < 0 || (B > 0 && (!nodes || !out))) { set_error("eu_get_node_type_host: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (B == 0) return EU_OK;
  unsigned long long* d_nodes = nullptr; int32_t* d_out = nullptr;
  EU_CUDA(cudaMalloc(&d_nodes, 8 * (size_t)B));
  cudaError_t e = cudaMalloc(&d_out, 4 * (size_t)B);
  int rc = EU_OK;
  if (e != cudaSuccess) rc = EU_ERR_CUDA;
  if (!rc && cudaMemcpyAsync(d_nodes, nodes, 8 * (size_t)B, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = EU_ERR_CUDA;
  if (!rc) rc = eu_get_node_type(c, (const int64_t*)d_nodes, B, d_out);
  if (!rc && (cudaMemcpyAsync(out, d_out, 4 * (size_t)B, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
              cudaStreamSynchronize(c->stream) != cudaSuccess)) rc = EU_ERR_CUDA;
  cudaFree(d_nodes); cudaFree(d_out);
  if (rc == EU_ERR_CUDA) set_error("eu_get_node_type_host: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}

extern "C" int eu_get_full_neighbor(eu_ctx* c, const int64_t* nodes, int64_t B, const int32_t* etypes, int32_t K,
                                    int64_t cap, int64_t* out_ptr, int64_t* out_ids, float* out_w, int32_t* out_t) {
  if (!c || B < 0 || K < 0 || K > EU_MAX_ETYPES || cap < 0 || !out_ptr || (B > 0 && !nodes) || (K > 0 && !etypes) ||
      (cap > 0 && (!out_ids || !out_w || !out_t))) {
    set_error("eu_get_full_neighbor: bad argument");
    return EU_ERR_INVALID;
  }
  EU_CUDA(cudaSetDevice(c->g->device));
  const DevGraph& d = c->g->d;
  ETList et{};
  et.K = K;
  for (int32_t k = 0; k < K; ++k) et.v[k] = etypes[k];
  cudaStream_t s = c->stream;
  size_t tmp = 0;
  cub::DeviceScan::InclusiveSum((void*)nullptr, tmp, (long long*)nullptr, (long long*)nullptr, (int)(B + 1), s);
  int rc = ctx_misc(c, (int64_t)tmp + 256);
  if (rc) return rc;
  { EuProfScope ps(c, "k_full_len", B);
    k_full_len<<<(unsigned)ceil_div(std::max<int64_t>(B, 1), 256), 256, 0, s>>>(d, (const unsigned long long*)nodes, B, et, (long long*)out_ptr); }
  EU_LAUNCHED();
  EU_CUDA(cub::DeviceScan::InclusiveSum(c->d_misc, tmp, (long long*)out_ptr, (long long*)out_ptr, (int)(B + 1), s));
  EU_LAUNCHED();
  if (cap > 0 && B > 0) {
    EuProfScope ps(c, "k_full_fill", B);
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(B * 32, 256), 148 * 8);
    k_full_fill<<<blocks, 256, 0, s>>>(d, (const unsigned long long*)nodes, B, et, (const long long*)out_ptr, cap,
                                       (unsigned long long*)out_ids, out_w, out_t);
    EU_LAUNCHED();
  }
  return EU_OK;
}

// Host buffers.  Call with cap = 0 to learn *total (out_ptr is filled), then with cap >= *total for the entries.
extern "C" int eu_get_full_neighbor_host(eu_ctx* c, const int64_t* nodes, int64_t B, const int32_t* etypes, int32_t K,
                                         int64_t cap, int64_t* out_ptr, int64_t* out_ids, float* out_w, int32_t* out_t,
                                         int64_t* total) {
  if (!c || B < 0 || cap < 0 || !out_ptr || (B > 0 && !nodes)) { set_error("eu_get_full_neighbor_host: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  unsigned long long* d_nodes = nullptr;
  long long* d_ptr = nullptr;
  EU_CUDA(cudaMalloc(&d_nodes, 8 * (size_t)std::max<int64_t>(B, 1)));
  EU_CUDA(cudaMalloc(&d_ptr, 8 * (size_t)(B + 1)));
  int rc = EU_OK;
  unsigned long long* d_ids = nullptr; float* d_w = nullptr; int32_t* d_t = nullptr;
  do {
    if (B > 0 && cudaMemcpyAsync(d_nodes, nodes, 8 * (size_t)B, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = EU_ERR_CUDA; break; }
    rc = eu_get_full_neighbor(c, (const int64_t*)d_nodes, B, etypes, K, 0, (int64_t*)d_ptr, nullptr, nullptr, nullptr);
    if (rc) break;
    if (cudaMemcpyAsync(out_ptr, d_ptr, 8 * (size_t)(B + 1), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = EU_ERR_CUDA; break; }
    const int64_t tot = out_ptr[B];
    if (total) *total = tot;
    const int64_t n = std::min(cap, tot);
    if (n > 0) {
      if (!out_ids || !out_w || !out_t) { set_error("eu_get_full_neighbor_host: null output"); rc = EU_ERR_INVALID; break; }
      if (cudaMalloc(&d_ids, 8 * (size_t)n) != cudaSuccess || cudaMalloc(&d_w, 4 * (size_t)n) != cudaSuccess ||
          cudaMalloc(&d_t, 4 * (size_t)n) != cudaSuccess) { set_error("eu_get_full_neighbor_host: cudaMalloc failed"); rc = EU_ERR_CUDA; break; }
      rc = eu_get_full_neighbor(c, (const int64_t*)d_nodes, B, etypes, K, n, (int64_t*)d_ptr, (int64_t*)d_ids, d_w, d_t);
      if (rc) break;
      cudaMemcpyAsync(out_ids, d_ids, 8 * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
      cudaMemcpyAsync(out_w, d_w, 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
      cudaMemcpyAsync(out_t, d_t, 4 * (size_t)n, cudaMemcpyDeviceToHost, c->stream);
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = EU_ERR_CUDA; break; }
    }
  } while (false);
  cudaFree(d_nodes); cudaFree(d_ptr); cudaFree(d_ids); cudaFree(d_w); cudaFree(d_t);
  if (rc == EU_ERR_CUDA) set_error("eu_get_full_neighbor_host: CUDA error %s", cudaGetErrorString(cudaGetLastError()));
  return rc;
}
