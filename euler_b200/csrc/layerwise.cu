// Layer-wise neighbor sampling and batch adjacency: FastGCN / AS-GCN / LGCN style dataflows.
// Reference semantics (file:line relative to /root/reference):
//   tf_euler SampleNeighborLayerwiseWithAdj   tf_euler/kernels/sample_neighbor_layerwise_with_adj_op.cc:54-150
//       query v(nodes).sampleLNB(edge_types, n, m, [weight_func,] default_node): per batch row of n nodes, m neighbors drawn from
//       the UNION of their neighbor lists; adj[b, j, k] = 1 iff out[b, k] is a neighbor of nodes[b, j]
//   API_LOCAL_SAMPLE_L                        euler/core/kernels/local_sample_layer_op.cc:41-140
//       candidates = the batch's full neighbors made unique by (dst, type) with their weights SUMMED in listing order, optional
//       sqrt, CompactWeightedCollection over them (sequential f32 prefix), m draws of one uniform each; an empty / zero-weight
//       candidate set fills default_node
//   API_SPARSE_GET_ADJ / tf_euler SparseGetAdj euler/core/kernels/sparse_get_adj_op.cc:34-90, tf_euler/kernels/sparse_get_adj_op.cc
//       adj[b, j, k] = 1 iff an edge (nodes[b, j], nb[b, k], t) exists for a listed type t
// Parity note: the reference enumerates the candidates in the iteration order of a std::unordered_map<std::string, ...> keyed
// by to_string(dst) + to_string(type) (local_sample_layer_op.cc:77-90) -- an order with no meaning that its own tests do not
// pin (neighbor_ops_test.py:142-181 check membership and adj consistency only).  Here the candidates are ordered by
// (dst, type): the candidate SET, every candidate's summed weight (same additions, same order) and hence the sampling
// DISTRIBUTION are identical; which uniform maps to which candidate differs.  tests/ check exactly that.
#include <cub/device/device_segmented_sort.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>

#include "internal.h"

namespace eu {

__global__ void k_lw_iota(long long* a, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a[i] = i;
}
__global__ void k_lw_gather_type(const long long* __restrict__ idx, const int32_t* __restrict__ t, int64_t n, int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = t[idx[i]];
}
__global__ void k_lw_gather_id(const long long* __restrict__ idx, const unsigned long long* __restrict__ ids, int64_t n,
                               unsigned long long* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = ids[idx[i]];
}
// segment b = the listing entries of batch row b: [ptr[b*n], ptr[(b+1)*n])
struct BatchSeg {
  const long long* ptr; long long n;
  __host__ __device__ long long operator()(long long b) const { return ptr[b * n]; }
};

// One thread per batch row walks its (dst, type)-sorted entries: unique candidates, weights summed in listing order (the
// sorts are stable), optional sqrt, sequential f32 prefix (CompactWeightedCollection::Init, compact_weighted_collection.h:82-97)
// written in place over the entry arrays.  cand_n[b] = candidates, cand_total[b] = their prefix end.
__global__ void k_lw_candidates(const long long* __restrict__ ptr, int64_t batch, int32_t n, const long long* __restrict__ order,
                                const unsigned long long* __restrict__ ids, const float* __restrict__ w, const int32_t* __restrict__ t,
                                int32_t weight_func, unsigned long long* __restrict__ c_id, int32_t* __restrict__ c_t,
                                float* __restrict__ c_cum, int32_t* __restrict__ cand_n, float* __restrict__ cand_total) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const long long lo = ptr[b * n], hi = ptr[(b + 1) * (int64_t)n];
  long long o = lo;       // candidates of this batch row are compacted to [lo, lo + cand_n)
  float run = 0.f;
  long long k = lo;
  while (k < hi) {
    const long long e0 = order[k];
    const unsigned long long id = ids[e0];
    const int32_t ty = t[e0];
    float sum = w[e0];
    long long k2 = k + 1;
    while (k2 < hi) {
      const long long e = order[k2];
      if (ids[e] != id || t[e] != ty) break;
      sum = __fadd_rn(sum, w[e]);
      ++k2;
    }
    if (weight_func == 1) sum = sqrtf(sum);          // local_sample_layer_op.cc:93-101 (float sqrt)
    run = __fadd_rn(run, sum);
    c_id[o] = id; c_t[o] = ty; c_cum[o] = run;
    ++o;
    k = k2;
  }
  cand_n[b] = (int32_t)(o - lo);
  cand_total[b] = run;
}

// serial prefix over the batch rows: how many uniforms the rows before b consumed (count each, only rows that sample)
__global__ void k_lw_positions(const int32_t* __restrict__ cand_n, const float* __restrict__ cand_total, int64_t batch, int32_t count,
                               unsigned long long* __restrict__ pos, EuRngState* rng, bool minstd) {
  if (blockIdx.x || threadIdx.x) return;
  unsigned long long run = 0;
  for (int64_t b = 0; b < batch; ++b) {
    pos[b] = run;
    if (cand_n[b] > 0 && cand_total[b] != 0.f) run += (unsigned long long)count;
  }
  pos[batch] = run;
  if (minstd) { rng->x_prev = rng->x; rng->x = modmul(rng->x, modpow_a(2ull * run)); rng->draws += run; }
  rng->calls += 1;
}

__global__ void k_lw_sample(const long long* __restrict__ ptr, int64_t batch, int32_t n, int32_t count, long long default_node,
                            const unsigned long long* __restrict__ c_id, const float* __restrict__ c_cum, const int32_t* __restrict__ cand_n,
                            const float* __restrict__ cand_total, const unsigned long long* __restrict__ pos, const EuRngState* rng,
                            bool philox, unsigned long long key, long long* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= batch * (int64_t)count) return;
  const int64_t b = i / count;
  const int32_t j = (int32_t)(i - b * count);
  const int32_t nc = cand_n[b];
  const float total = cand_total[b];
  if (nc == 0 || total == 0.f) { out[i] = default_node; return; }
  double u, u2;
  if (philox) philox_uniform2((unsigned long long)b, (uint32_t)j, (uint32_t)rng->calls, key ^ 0x4C57ull, u, u2);
  else { uint32_t x = modmul(rng->x_prev, modpow_a(2ull * (pos[b] + (unsigned long long)j))); u = minstd_uniform(x); }
  const float* cum = c_cum + ptr[b * n];
  const int32_t k = upper_bound_clamped(cum, 0, nc - 1, gt_threshold(pick_r(u, 0.f, total)));
  out[i] = (long long)c_id[ptr[b * n] + k];
}

// adj[b, j, k] = 1 iff nb[b, k] appears among the listed neighbors of nodes[b, j]; the listing slice of (b, j) is searched
// linearly (unsorted lists are legal)
__global__ void k_lw_adj(const long long* __restrict__ ptr, int64_t batch, int32_t n, int32_t m, const unsigned long long* __restrict__ ids,
                         const long long* __restrict__ nb, float* __restrict__ adj) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= batch * (int64_t)n * m) return;
  const int64_t bj = i / m;
  const int64_t b = bj / n;
  const int32_t k = (int32_t)(i - bj * m);
  const unsigned long long want = (unsigned long long)nb[b * m + k];
  float v = 0.f;
  for (long long e = ptr[bj]; e < ptr[bj + 1]; ++e)
    if (ids[e] == want) { v = 1.f; break; }
  adj[i] = v;
}

}  // namespace eu

using namespace eu;

// Shared front end: the full listing of batch * n nodes into scratch (synchronises once to size it).
struct LwListing {
  long long* ptr = nullptr; unsigned long long* ids = nullptr; float* w = nullptr; int32_t* t = nullptr; long long total = 0;
  void free_all() { cudaFree(ptr); cudaFree(ids); cudaFree(w); cudaFree(t); }
};
static int lw_listing(eu_ctx* c, const int64_t* nodes, int64_t rows, const int32_t* etypes, int32_t K, LwListing* L) {
  cudaStream_t s = c->stream;
  EU_CUDA(cudaMalloc(&L->ptr, 8 * (size_t)(rows + 1)));
  int rc = eu_get_full_neighbor(c, nodes, rows, etypes, K, 0, (int64_t*)L->ptr, nullptr, nullptr, nullptr);
  if (rc) return rc;
  EU_CUDA(cudaMemcpyAsync(&L->total, L->ptr + rows, 8, cudaMemcpyDeviceToHost, s));
  EU_CUDA(cudaStreamSynchronize(s));
  if (L->total >= ((long long)1 << 31)) { set_error("layerwise: more than 2^31 listed neighbors"); return EU_ERR_UNSUPPORTED; }
  const size_t nn = (size_t)std::max<long long>(L->total, 1);
  EU_CUDA(cudaMalloc(&L->ids, 8 * nn));
  EU_CUDA(cudaMalloc(&L->w, 4 * nn));
  EU_CUDA(cudaMalloc(&L->t, 4 * nn));
  if (L->total > 0) rc = eu_get_full_neighbor(c, nodes, rows, etypes, K, L->total, (int64_t*)L->ptr, (int64_t*)L->ids, L->w, L->t);
  return rc;
}

extern "C" int eu_sample_neighbor_layerwise(eu_ctx* c, const int64_t* nodes, int64_t batch, int32_t n, const int32_t* etypes, int32_t K,
                                            int32_t count, int64_t default_node, int32_t weight_func, int64_t* out_nb, float* out_adj) {
  if (!c || batch < 0 || n < 1 || count < 0 || weight_func < 0 || weight_func > 1 || (batch > 0 && (!nodes || (count > 0 && !out_nb)))) {
    set_error("eu_sample_neighbor_layerwise: bad argument");
    return EU_ERR_INVALID;
  }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (batch == 0 || count == 0) return EU_OK;
  if (batch >= ((int64_t)1 << 31)) { set_error("eu_sample_neighbor_layerwise: more than 2^31 batch rows"); return EU_ERR_UNSUPPORTED; }
  cudaStream_t s = c->stream;
  LwListing L;
  int rc = lw_listing(c, nodes, batch * n, etypes, K, &L);
  char* buf = nullptr;
  if (!rc) {
    const size_t nn = (size_t)std::max<long long>(L.total, 1);
    BatchSeg f{L.ptr, (long long)n};
    cub::CountingInputIterator<long long> cnt(0);
    cub::TransformInputIterator<long long, BatchSeg, cub::CountingInputIterator<long long>> seg(cnt, f);
    size_t tmp1 = 0, tmp2 = 0;
    cub::DeviceSegmentedSort::StableSortPairs((void*)nullptr, tmp1, (const int32_t*)nullptr, (int32_t*)nullptr, (const long long*)nullptr,
                                              (long long*)nullptr, (int)nn, (int)batch, seg, seg + 1, s);
    cub::DeviceSegmentedSort::StableSortPairs((void*)nullptr, tmp2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                              (const long long*)nullptr, (long long*)nullptr, (int)nn, (int)batch, seg, seg + 1, s);
    const size_t tmp = (std::max(tmp1, tmp2) + 255) & ~(size_t)255;
    // idx_a 8 | idx_b 8 | key64_a 8 | key64_b 8 | key32_a 4 | key32_b 4 | c_id 8 | c_cum 4 | c_t 4 | per batch: cand_n 4, total 4, pos 8(+1)
    const size_t per = 8 + 8 + 8 + 8 + 4 + 4 + 8 + 4 + 4;
    const size_t bytes = per * nn + 16 * (size_t)(batch + 2) + 4096 + tmp;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { set_error("eu_sample_neighbor_layerwise: cudaMalloc(%zu) failed", bytes); rc = EU_ERR_CUDA; }
    if (!rc) {
      char* p = buf;
      auto take = [&](size_t b) { char* q = p; p += (b + 255) & ~(size_t)255; return q; };
      long long* idx_a = (long long*)take(8 * nn); long long* idx_b = (long long*)take(8 * nn);
      unsigned long long* k64_a = (unsigned long long*)take(8 * nn); unsigned long long* k64_b = (unsigned long long*)take(8 * nn);
      int32_t* k32_a = (int32_t*)take(4 * nn); int32_t* k32_b = (int32_t*)take(4 * nn);
      unsigned long long* c_id = (unsigned long long*)take(8 * nn); float* c_cum = (float*)take(4 * nn); int32_t* c_t = (int32_t*)take(4 * nn);
      int32_t* cand_n = (int32_t*)take(4 * (size_t)batch); float* cand_total = (float*)take(4 * (size_t)batch);
      unsigned long long* pos = (unsigned long long*)take(8 * (size_t)(batch + 1));
      void* cubtmp = take(tmp);
      const long long* order = idx_a;
      if (L.total > 0) {
        // stable sort by type, then stable sort by neighbor id: entries ordered by (dst, type), equal keys in listing order
        k_lw_iota<<<148 * 4, 256, 0, s>>>(idx_a, L.total);
        cudaMemcpyAsync(k32_a, L.t, 4 * nn, cudaMemcpyDeviceToDevice, s);
        cudaError_t e = cub::DeviceSegmentedSort::StableSortPairs(cubtmp, tmp1, (const int32_t*)k32_a, k32_b, (const long long*)idx_a, idx_b,
                                                                  (int)L.total, (int)batch, seg, seg + 1, s);
        k_lw_gather_id<<<148 * 4, 256, 0, s>>>(idx_b, L.ids, L.total, k64_a);
        if (e == cudaSuccess)
          e = cub::DeviceSegmentedSort::StableSortPairs(cubtmp, tmp2, (const unsigned long long*)k64_a, k64_b, (const long long*)idx_b, idx_a,
                                                        (int)L.total, (int)batch, seg, seg + 1, s);
        g_launches += 4;
        if (e != cudaSuccess) { set_error("eu_sample_neighbor_layerwise: %s", cudaGetErrorString(e)); rc = EU_ERR_CUDA; }
      }
      if (!rc) {
        k_lw_candidates<<<(unsigned)ceil_div(batch, 128), 128, 0, s>>>(L.ptr, batch, n, order, L.ids, L.w, L.t, weight_func, c_id, c_t, c_cum,
                                                                      cand_n, cand_total);
        k_lw_positions<<<1, 32, 0, s>>>(cand_n, cand_total, batch, count, pos, c->d_rng, c->rng == EU_RNG_MINSTD);
        k_lw_sample<<<(unsigned)ceil_div(batch * (int64_t)count, 256), 256, 0, s>>>(L.ptr, batch, n, count, (long long)default_node, c_id, c_cum,
                                                                                   cand_n, cand_total, pos, c->d_rng, c->rng == EU_RNG_PHILOX,
                                                                                   c->seed, (long long*)out_nb);
        g_launches += 3;
        if (out_adj) {
          k_lw_adj<<<(unsigned)ceil_div(batch * (int64_t)n * count, 256), 256, 0, s>>>(L.ptr, batch, n, count, L.ids, (const long long*)out_nb, out_adj);
          g_launches++;
        }
        if (cudaStreamSynchronize(s) != cudaSuccess) { set_error("eu_sample_neighbor_layerwise: %s", cudaGetErrorString(cudaGetLastError())); rc = EU_ERR_CUDA; }
      }
    }
  }
  cudaFree(buf);
  L.free_all();
  return rc;
}

// tf_euler.sparse_get_adj: dense 0/1 view [batch, N, M] of the reference's SparseTensor
extern "C" int eu_sparse_get_adj(eu_ctx* c, const int64_t* nodes, const int64_t* nb_nodes, int64_t batch, int32_t N, int32_t M,
                                 const int32_t* etypes, int32_t K, float* out_adj) {
  if (!c || batch < 0 || N < 1 || M < 1 || (batch > 0 && (!nodes || !nb_nodes || !out_adj))) { set_error("eu_sparse_get_adj: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (batch == 0) return EU_OK;
  LwListing L;
  int rc = lw_listing(c, nodes, batch * N, etypes, K, &L);
  if (!rc) {
    k_lw_adj<<<(unsigned)ceil_div(batch * (int64_t)N * M, 256), 256, 0, c->stream>>>(L.ptr, batch, N, M, L.ids, (const long long*)nb_nodes, out_adj);
    g_launches++;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) { set_error("eu_sparse_get_adj: %s", cudaGetErrorString(cudaGetLastError())); rc = EU_ERR_CUDA; }
  }
  L.free_all();
  return rc;
}
