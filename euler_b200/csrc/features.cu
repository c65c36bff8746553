// Ragged node features on the HBM-resident graph: uint64 ("sparse") and binary features.
// Reference semantics (file:line relative to /root/reference):
//   Node::GetUint64Feature / GetBinaryFeature   euler/core/graph/node.cc:330-409   slot s = [idx[s-1], idx[s]) of the node's array
//   tf_euler GetSparseFeature                   tf_euler/kernels/get_sparse_feature_op.cc:52-130  a node without values gets one
//                                               entry {i, 0} = default_value
//   tf_euler GetBinaryFeature                   tf_euler/kernels/get_binary_feature_op.cc          one string per node
// Same shape as the full-neighbor listing: per-node lengths -> cub inclusive scan -> copy; no host sync in the device entry points.
#include <cub/device/device_scan.cuh>

#include <algorithm>

#include "internal.h"

namespace eu {

// row slice of slot `fid`: [b, e) in the value array, b == e when the node / slot does not exist
__device__ __forceinline__ void ragged_slice(const int64_t* __restrict__ ptr, int32_t S, int64_t row, int32_t fid, int64_t* b, int64_t* e) {
  *b = *e = 0;
  if (row < 0 || fid < 0 || fid >= S || !ptr) return;
  *b = ptr[row * S + fid];
  *e = ptr[row * S + fid + 1];
}

template <bool SPARSE>
__global__ void k_ragged_len(DevGraph g, const unsigned long long* __restrict__ nodes, int64_t M, int32_t fid,
                             long long* __restrict__ out_ptr) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i == 0) out_ptr[0] = 0;
  if (i >= M) return;
  int64_t b, e;
  const int64_t row = lookup_row(g, nodes[i]);
  if (SPARSE) ragged_slice(g.u64_ptr, g.n_u64_slots, row, fid, &b, &e);
  else ragged_slice(g.bin_ptr, g.n_bin_slots, row, fid, &b, &e);
  long long len = e - b;
  if (SPARSE && len == 0) len = 1;   // one default entry (get_sparse_feature_op.cc:96-99)
  out_ptr[i + 1] = len;
}

template <bool SPARSE>
__global__ void __launch_bounds__(256) k_ragged_fill(DevGraph g, const unsigned long long* __restrict__ nodes, int64_t M, int32_t fid,
                                                     long long default_value, const long long* __restrict__ out_ptr, int64_t cap,
                                                     long long* __restrict__ out_values, unsigned char* __restrict__ out_bytes) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; i < M; i += nwarps) {
    int64_t b, e;
    const int64_t row = lookup_row(g, nodes[i]);
    if (SPARSE) ragged_slice(g.u64_ptr, g.n_u64_slots, row, fid, &b, &e);
    else ragged_slice(g.bin_ptr, g.n_bin_slots, row, fid, &b, &e);
    const int64_t o = out_ptr[i];
    if (SPARSE) {
      if (e == b) { if (lane == 0 && o < cap) out_values[o] = default_value; continue; }
      for (int64_t k = lane; k < e - b; k += 32) if (o + k < cap) out_values[o + k] = (long long)g.u64_val[b + k];
    } else {
      for (int64_t k = lane; k < e - b; k += 32) if (o + k < cap) out_bytes[o + k] = g.bin_val[b + k];
    }
  }
}

template <bool SPARSE>
static int ragged_get(eu_ctx* c, const int64_t* nodes, int64_t M, int32_t fid, int64_t default_value, int64_t cap, int64_t* out_ptr,
                      int64_t* out_values, uint8_t* out_bytes, const char* what) {
  if (!c || M < 0 || cap < 0 || !out_ptr || (M > 0 && !nodes) || (cap > 0 && !(SPARSE ? (void*)out_values : (void*)out_bytes))) {
    set_error("%s: bad argument", what);
    return EU_ERR_INVALID;
  }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (M >= ((int64_t)1 << 31)) { set_error("%s: more than 2^31 nodes", what); return EU_ERR_UNSUPPORTED; }
  const DevGraph& d = c->g->d;
  cudaStream_t s = c->stream;
  size_t tmp = 0;
  cub::DeviceScan::InclusiveSum((void*)nullptr, tmp, (long long*)nullptr, (long long*)nullptr, (int)(M + 1), s);
  int rc = ctx_misc(c, (int64_t)tmp + 256);
  if (rc) return rc;
  k_ragged_len<SPARSE><<<(unsigned)ceil_div(std::max<int64_t>(M, 1), 256), 256, 0, s>>>(d, (const unsigned long long*)nodes, M, fid, (long long*)out_ptr);
  EU_LAUNCHED();
  EU_CUDA(cub::DeviceScan::InclusiveSum(c->d_misc, tmp, (long long*)out_ptr, (long long*)out_ptr, (int)(M + 1), s));
  EU_LAUNCHED();
  if (cap > 0 && M > 0) {
    const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(M * 32, 256), 148 * 8);
    k_ragged_fill<SPARSE><<<blocks, 256, 0, s>>>(d, (const unsigned long long*)nodes, M, fid, (long long)default_value, (const long long*)out_ptr, cap,
                                                 (long long*)out_values, out_bytes);
    EU_LAUNCHED();
  }
  return EU_OK;
}

// host buffers: lengths first (total), then the values when cap allows
template <bool SPARSE>
static int ragged_get_host(eu_ctx* c, const int64_t* nodes, int64_t M, int32_t fid, int64_t default_value, int64_t cap, int64_t* out_ptr,
                           void* out_vals, int64_t* total, const char* what) {
  if (!c || M < 0 || cap < 0 || !out_ptr || (M > 0 && !nodes)) { set_error("%s: bad argument", what); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  unsigned long long* d_nodes = nullptr;
  long long* d_ptr = nullptr;
  char* d_vals = nullptr;
  EU_CUDA(cudaMalloc(&d_nodes, 8 * (size_t)std::max<int64_t>(M, 1)));
  EU_CUDA(cudaMalloc(&d_ptr, 8 * (size_t)(M + 1)));
  const size_t esz = SPARSE ? 8 : 1;
  int rc = EU_OK;
  do {
    if (M > 0 && cudaMemcpyAsync(d_nodes, nodes, 8 * (size_t)M, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) { rc = EU_ERR_CUDA; break; }
    rc = ragged_get<SPARSE>(c, (const int64_t*)d_nodes, M, fid, default_value, 0, (int64_t*)d_ptr, nullptr, nullptr, what);
    if (rc) break;
    if (cudaMemcpyAsync(out_ptr, d_ptr, 8 * (size_t)(M + 1), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
        cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = EU_ERR_CUDA; break; }
    const int64_t tot = out_ptr[M];
    if (total) *total = tot;
    const int64_t n = std::min(cap, tot);
    if (n > 0) {
      if (!out_vals) { set_error("%s: null output", what); rc = EU_ERR_INVALID; break; }
      if (cudaMalloc(&d_vals, esz * (size_t)n) != cudaSuccess) { set_error("%s: cudaMalloc failed", what); rc = EU_ERR_CUDA; break; }
      rc = ragged_get<SPARSE>(c, (const int64_t*)d_nodes, M, fid, default_value, n, (int64_t*)d_ptr, SPARSE ? (int64_t*)d_vals : nullptr,
                              SPARSE ? nullptr : (uint8_t*)d_vals, what);
      if (rc) break;
      if (cudaMemcpyAsync(out_vals, d_vals, esz * (size_t)n, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
          cudaStreamSynchronize(c->stream) != cudaSuccess) { rc = EU_ERR_CUDA; break; }
    }
  } while (false);
  cudaFree(d_nodes); cudaFree(d_ptr); cudaFree(d_vals);
  if (rc == EU_ERR_CUDA) set_error("%s: CUDA error %s", what, cudaGetErrorString(cudaGetLastError()));
  return rc;
}

}  // namespace eu

using namespace eu;

extern "C" {

int eu_get_sparse_feature(eu_ctx* c, const int64_t* nodes, int64_t M, int32_t fid, int64_t default_value, int64_t cap, int64_t* out_ptr,
                          int64_t* out_values) {
  return ragged_get<true>(c, nodes, M, fid, default_value, cap, out_ptr, out_values, nullptr, "eu_get_sparse_feature");
}
int eu_get_sparse_feature_host(eu_ctx* c, const int64_t* nodes, int64_t M, int32_t fid, int64_t default_value, int64_t cap, int64_t* out_ptr,
                               int64_t* out_values, int64_t* total) {
  return ragged_get_host<true>(c, nodes, M, fid, default_value, cap, out_ptr, out_values, total, "eu_get_sparse_feature_host");
}
int eu_get_binary_feature(eu_ctx* c, const int64_t* nodes, int64_t M, int32_t fid, int64_t cap, int64_t* out_ptr, uint8_t* out_bytes) {
  return ragged_get<false>(c, nodes, M, fid, 0, cap, out_ptr, nullptr, out_bytes, "eu_get_binary_feature");
}
int eu_get_binary_feature_host(eu_ctx* c, const int64_t* nodes, int64_t M, int32_t fid, int64_t cap, int64_t* out_ptr, uint8_t* out_bytes,
                               int64_t* total) {
  return ragged_get_host<false>(c, nodes, M, fid, 0, cap, out_ptr, out_bytes, total, "eu_get_binary_feature_host");
}

}  // extern "C"
