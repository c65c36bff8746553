// Peer-memory exchange for the sharded path: the all-to-all of a hop / feature fetch is done by the kernels
// themselves with loads/stores on NVLink peer mappings -- no NCCL call, no host-visible counts, no host sync,
// CUDA-graph capturable.  Replaces euler's ID_SPLIT -> REMOTE(gRPC) -> IDX_MERGE/DATA_MERGE
// (euler/core/kernels/id_split_op.cc:46-99, remote_op.cc:60-146, idx_merge_op.cc:32-78).
//
// Every rank owns one symmetric region (same layout everywhere, cudaIpc-mapped into every peer):
//   header | inbox_ids[N][cap] | inbox_src[N][cap] | out_eng | out_ids | out_w | out_t | out_rows
// One exchange =
//   requester  k_sym_push      : its bucket for owner o goes straight into o's inbox segment [me] (+ the original row
//                                index of every seed), then count + flagA[me] <- epoch on o
//   owner      k_sym_wait_pad  : waits for flagA of every source, zero-pads each segment (id 0 = "exists nowhere")
//   owner      hop() / gather  : samples the padded inbox as ONE sampleNB call on its own engine (sources in rank order,
//                                each in batch order: the order pinned in euler_b200/sharded.py) -- or, for features,
//   owner      k_sym_reply_*   : writes every result row DIRECTLY into the requester's output arrays at the seed's
//                                original position (TF packing done here), then flagB[me] <- epoch on the requester
//   requester  k_sym_wait      : waits for flagB of every owner; its outputs are complete, already in request order.
// Flags are monotonic epochs kept on the device; all waits are bounded (hdr->error is set on timeout instead of hanging).
#include <string.h>

#include <algorithm>

#include "internal.h"

namespace eu {

static constexpr int kSymMaxRanks = 16;

struct SymHeader {
  unsigned int flagA[kSymMaxRanks];   // [src]   epoch of the last inbox segment pushed by src
  unsigned int flagB[kSymMaxRanks];   // [owner] epoch of the last reply written by owner
  int in_cnt[kSymMaxRanks];           // [src]   seeds in src's segment
  unsigned int epoch;                 // local exchange counter
  unsigned int done;                  // last-block ticket
  int error;                          // 1 = a wait timed out
  int pad;
};

struct SymLayout {
  int64_t cap;          // inbox slots per source
  int64_t max_out;      // rows * count slots of the sample outputs
  int64_t max_rows_f;   // rows of the feature output
  int32_t max_dim;
  int64_t off_inbox_ids, off_inbox_src, off_eng, off_ids, off_w, off_t, off_rows, bytes;
};

struct SymPeers {
  char* base[kSymMaxRanks];
};

__device__ __forceinline__ SymHeader* hdr_of(char* base) { return reinterpret_cast<SymHeader*>(base); }

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// bounded spin until *flag >= want
__device__ __forceinline__ bool spin_until(const unsigned int* flag, unsigned int want, int* error) {
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys(flag) - want) < 0) {
    __nanosleep(200);
    if (clock64() - t0 > 4000000000LL) { *error = 1; return false; }   // ~2 s
  }
  return true;
}

// ---- requester: push my sorted bucket segments into the owners' inboxes
__global__ void __launch_bounds__(256) k_sym_push(SymPeers peers, SymLayout lay, int me, int N,
                                                  const unsigned long long* __restrict__ sorted_ids,
                                                  const int32_t* __restrict__ src_index,
                                                  const long long* __restrict__ offsets /*[N+1]*/, int64_t rows) {
  __shared__ long long s_off[kSymMaxRanks + 1];
  __shared__ bool s_last;
  if (threadIdx.x <= N) s_off[threadIdx.x] = offsets[threadIdx.x];
  __syncthreads();
  // persistent grid: one system fence per block, not per 256 rows (a fence waits for the NVLink acks of the block's stores)
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < rows; k += (int64_t)gridDim.x * blockDim.x) {
    int o = 0;
    while (o + 1 < N && k >= s_off[o + 1]) ++o;
    const int64_t pos = k - s_off[o];
    char* pb = peers.base[o];
    reinterpret_cast<unsigned long long*>(pb + lay.off_inbox_ids)[(int64_t)me * lay.cap + pos] = sorted_ids[k];
    reinterpret_cast<int32_t*>(pb + lay.off_inbox_src)[(int64_t)me * lay.cap + pos] = src_index[k];
  }
  __threadfence_system();
  __syncthreads();
  SymHeader* mine = hdr_of(peers.base[me]);
  if (threadIdx.x == 0) s_last = atomicAdd(&mine->done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  unsigned int e = 0;
  if (threadIdx.x == 0) { e = mine->epoch + 1; mine->epoch = e; mine->done = 0; }
  __shared__ unsigned int s_e;
  if (threadIdx.x == 0) s_e = e;
  __syncthreads();
  e = s_e;
  if (threadIdx.x < N) {
    SymHeader* h = hdr_of(peers.base[threadIdx.x]);
    h->in_cnt[me] = (int)(s_off[threadIdx.x + 1] - s_off[threadIdx.x]);
    __threadfence_system();
    st_release_sys(&h->flagA[me], e);
  }
}

// ---- owner: wait for every source (one small block spins; nothing else of the GPU is held)
__global__ void k_sym_wait_in(char* base, int N) {
  SymHeader* h = hdr_of(base);
  if (threadIdx.x < N) spin_until(&h->flagA[threadIdx.x], h->epoch, &h->error);
}

// ---- owner: zero-pad the segments (the counts were published before the flags)
__global__ void __launch_bounds__(256) k_sym_wait_pad(char* base, SymLayout lay, int N) {
  SymHeader* h = hdr_of(base);
  __shared__ int s_cnt[kSymMaxRanks];
  if (threadIdx.x < N) s_cnt[threadIdx.x] = *reinterpret_cast<volatile int*>(&h->in_cnt[threadIdx.x]);
  __syncthreads();
  unsigned long long* ids = reinterpret_cast<unsigned long long*>(base + lay.off_inbox_ids);
  const int64_t total = (int64_t)N * lay.cap;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(i / lay.cap);
    if (i - s * lay.cap >= s_cnt[s]) ids[i] = 0ull;
  }
}

// ---- owner: sampled rows of the padded inbox -> requester's outputs at the original positions (+ TF packing)
__global__ void __launch_bounds__(256) k_sym_reply_sample(SymPeers peers, SymLayout lay, int me, int N, int32_t count,
                                                          long long default_node, const long long* __restrict__ r_ids,
                                                          const float* __restrict__ r_w, const int32_t* __restrict__ r_t,
                                                          bool want_packed) {
  char* base = peers.base[me];
  SymHeader* mine = hdr_of(base);
  const int32_t* src = reinterpret_cast<const int32_t*>(base + lay.off_inbox_src);
  __shared__ bool s_last;
  const int64_t slots = (int64_t)N * lay.cap * count;
  for (int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; tid < slots; tid += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = tid / count;            // padded inbox row
    const int32_t j = (int32_t)(tid - row * count);
    const int s = (int)(row / lay.cap);
    const int64_t k = row - (int64_t)s * lay.cap;
    if (k < mine->in_cnt[s]) {
      const int64_t dst = (int64_t)src[row] * count + j;
      const long long id = r_ids[tid];
      const bool keep = r_ids[row * count] != 0;   // tf_euler/kernels/sample_neighbor_op.cc:114-122
      char* pb = peers.base[s];
      reinterpret_cast<long long*>(pb + lay.off_eng)[dst] = id;
      if (want_packed) {
        reinterpret_cast<long long*>(pb + lay.off_ids)[dst] = keep ? id : default_node;
        reinterpret_cast<float*>(pb + lay.off_w)[dst] = keep ? r_w[tid] : 0.f;
        reinterpret_cast<int32_t*>(pb + lay.off_t)[dst] = keep ? r_t[tid] : -1;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&mine->done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) mine->done = 0;
  if (threadIdx.x < N) st_release_sys(&hdr_of(peers.base[threadIdx.x])->flagB[me], mine->epoch);
}

// ---- owner: feature rows gathered from the local shard straight into the requester's output (G lanes per row)
__global__ void __launch_bounds__(256) k_sym_reply_feature(DevGraph g, SymPeers peers, SymLayout lay, int me, int N,
                                                           int32_t dim, int32_t soff, int32_t sdim, int G) {
  char* base = peers.base[me];
  SymHeader* mine = hdr_of(base);
  const unsigned long long* ids = reinterpret_cast<const unsigned long long*>(base + lay.off_inbox_ids);
  const int32_t* src = reinterpret_cast<const int32_t*>(base + lay.off_inbox_src);
  __shared__ bool s_last;
  const int sh = 31 - __clz(G);
  const int sub = (int)(threadIdx.x & (G - 1));
  for (int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> sh; row < (int64_t)N * lay.cap;
       row += ((int64_t)gridDim.x * blockDim.x) >> sh) {
    const int s = (int)(row / lay.cap);
    if (row - (int64_t)s * lay.cap < mine->in_cnt[s]) {
      const int64_t gr = sdim > 0 ? lookup_row(g, ids[row]) : -1;
      const float* f = gr >= 0 ? g.feat + gr * (int64_t)g.feat_dim + soff : nullptr;
      float* o = reinterpret_cast<float*>(peers.base[s] + lay.off_rows) + (int64_t)src[row] * dim;
      for (int32_t d = sub * 4; d < dim; d += G * 4) {   // dim, sdim, soff multiples of 4 (checked by the launcher)
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f && d < sdim) v = __ldg(reinterpret_cast<const float4*>(f + d));
        *reinterpret_cast<float4*>(o + d) = v;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&mine->done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) mine->done = 0;
  if (threadIdx.x < N) st_release_sys(&hdr_of(peers.base[threadIdx.x])->flagB[me], mine->epoch);
}

// ---- requester: wait for every owner's replies
__global__ void k_sym_wait(char* base, int N) {
  SymHeader* h = hdr_of(base);
  if (threadIdx.x < N) spin_until(&h->flagB[threadIdx.x], h->epoch, &h->error);
}

// ---- fused sharded SAGE mean.  The requester pushes the neighbor ids of its fixed-fanout block (flat [rows*count],
// src index = flat position, stable bucket => ascending inside every segment); the owner sums the features of ITS ids
// per destination row, j ascending, and stores ONE partial row per (owner, destination) into the requester's `part`
// region; the requester adds the N partials in rank order and divides.  NVLink carries rows*dim floats per owner instead
// of rows*count*dim: the hop-2 features never cross the link (sage_dataflow.py:43-46 + mp_ops.py:65-69 fused with the
// REMOTE get_dense_feature they follow).
__device__ __forceinline__ int32_t warp_lower_bound(const int32_t* __restrict__ a, int32_t n, int32_t key, int lane) {
  int32_t lo = 0, hi = n;
  while (hi - lo > 32) {
    const int32_t len = hi - lo;
    const int32_t p = lo + (int32_t)(((int64_t)(lane + 1) * len) / 33);   // lo < p_0 <= .. <= p_31 < hi
    const int c = __popc(__ballot_sync(0xffffffffu, a[p] < key));          // sorted: a prefix of the probes is below
    const int32_t below = __shfl_sync(0xffffffffu, p, c > 0 ? c - 1 : 0);
    const int32_t above = __shfl_sync(0xffffffffu, p, c < 32 ? c : 31);
    lo = c > 0 ? below + 1 : lo;
    hi = c < 32 ? above : hi;
  }
  const bool b = lo + lane < hi && a[lo + lane] < key;
  return lo + __popc(__ballot_sync(0xffffffffu, b));
}

template <int NV>   // feat_dim == dim == NV*128
__global__ void __launch_bounds__(256) k_sym_reply_sage(DevGraph g, SymPeers peers, SymLayout lay, int me, int N, int64_t rows,
                                                        int32_t count) {
  char* base = peers.base[me];
  SymHeader* mine = hdr_of(base);
  const unsigned long long* ids = reinterpret_cast<const unsigned long long*>(base + lay.off_inbox_ids);
  const int32_t* src = reinterpret_cast<const int32_t*>(base + lay.off_inbox_src);
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31;
  constexpr int32_t fd = NV * 128;
  const float* __restrict__ feat = g.feat + lane * 4;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; w < (int64_t)N * rows; w += nwarps) {
    const int s = (int)(w / rows);
    const int64_t d = w - (int64_t)s * rows;
    const int32_t n_s = mine->in_cnt[s];
    const int32_t* seg = src + (int64_t)s * lay.cap;
    const unsigned long long* sid = ids + (int64_t)s * lay.cap;
    const int32_t key_hi = (int32_t)((d + 1) * count);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t e0 = warp_lower_bound(seg, n_s, (int32_t)(d * count), lane);; e0 += 32) {
      const int32_t e = e0 + lane;
      const bool in = e < n_s && seg[e] < key_hi;
      int32_t my = -1;
      if (in) my = (int32_t)lookup_row(g, sid[e]);
      const int nin = __popc(__ballot_sync(0xffffffffu, in));
      unsigned valid = __ballot_sync(0xffffffffu, my >= 0);
      while (valid) {
        float4 v[4][NV];
        int n = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (valid) {
            const int j = __ffs(valid) - 1;
            valid &= valid - 1;
            const int32_t row = __shfl_sync(0xffffffffu, my, j);
            const float* p = feat + (int64_t)row * fd;
#pragma unroll
            for (int t = 0; t < NV; ++t) v[q][t] = __ldg(reinterpret_cast<const float4*>(p + t * 128));
            n = q + 1;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (q < n) {
#pragma unroll
            for (int t = 0; t < NV; ++t) {
              acc[t].x = __fadd_rn(acc[t].x, v[q][t].x); acc[t].y = __fadd_rn(acc[t].y, v[q][t].y);
              acc[t].z = __fadd_rn(acc[t].z, v[q][t].z); acc[t].w = __fadd_rn(acc[t].w, v[q][t].w);
            }
          }
        }
      }
      if (nin < 32) break;
    }
    float* o = reinterpret_cast<float*>(peers.base[s] + lay.off_rows) + ((int64_t)me * rows + d) * fd + lane * 4;
#pragma unroll
    for (int t = 0; t < NV; ++t) *reinterpret_cast<float4*>(o + t * 128) = acc[t];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&mine->done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) mine->done = 0;
  if (threadIdx.x < N) st_release_sys(&hdr_of(peers.base[threadIdx.x])->flagB[me], mine->epoch);
}

// any width: lanes over columns, entries serial (slow path, same sums)
__global__ void __launch_bounds__(256) k_sym_reply_sage_generic(DevGraph g, SymPeers peers, SymLayout lay, int me, int N, int64_t rows,
                                                                int32_t count, int32_t dim) {
  char* base = peers.base[me];
  SymHeader* mine = hdr_of(base);
  const unsigned long long* ids = reinterpret_cast<const unsigned long long*>(base + lay.off_inbox_ids);
  const int32_t* src = reinterpret_cast<const int32_t*>(base + lay.off_inbox_src);
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31;
  const int32_t fd = g.feat_dim;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; w < (int64_t)N * rows; w += nwarps) {
    const int s = (int)(w / rows);
    const int64_t d = w - (int64_t)s * rows;
    const int32_t n_s = mine->in_cnt[s];
    const int32_t* seg = src + (int64_t)s * lay.cap;
    const unsigned long long* sid = ids + (int64_t)s * lay.cap;
    const int32_t key_hi = (int32_t)((d + 1) * count);
    const int32_t e_lo = warp_lower_bound(seg, n_s, (int32_t)(d * count), lane);
    float* o = reinterpret_cast<float*>(peers.base[s] + lay.off_rows) + ((int64_t)me * rows + d) * dim;
    for (int32_t c0 = 0; c0 < dim; c0 += 32) {
      const int32_t col = c0 + lane;
      float acc = 0.f;
      for (int32_t e = e_lo; e < n_s && seg[e] < key_hi; ++e) {   // warp-uniform
        const int64_t row = lookup_row(g, sid[e]);
        if (row >= 0 && col < fd && col < dim) acc = __fadd_rn(acc, __ldg(g.feat + row * (int64_t)fd + col));
      }
      if (col < dim) o[col] = acc;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(&mine->done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  if (threadIdx.x == 0) mine->done = 0;
  if (threadIdx.x < N) st_release_sys(&hdr_of(peers.base[threadIdx.x])->flagB[me], mine->epoch);
}

// requester: out[d,:] = (part[0][d,:] + part[1][d,:] + ...) / (count + 1e-7), rank order
__global__ void __launch_bounds__(256) k_sym_sage_reduce(const float* __restrict__ part, int N, int64_t rows, int32_t dim, int32_t count,
                                                         float* __restrict__ out) {
  const float denom = __fadd_rn((float)count, 1e-7f);
  const int64_t total = rows * dim;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = part[i];
    for (int o = 1; o < N; ++o) acc = __fadd_rn(acc, part[(int64_t)o * total + i]);
    out[i] = __fdiv_rn(acc, denom);
  }
}

static inline unsigned sym_grid(int64_t threads) {   // persistent: at most 8 CTAs of 256 per SM
  return (unsigned)std::min<int64_t>(std::max<int64_t>(ceil_div(threads, 256), 1), 148 * 8);
}

static inline int64_t a256(int64_t x) { return (x + 255) & ~(int64_t)255; }

}  // namespace eu

struct eu_sym {
  eu_ctx* c = nullptr;
  int rank = 0, world = 1;
  eu::SymLayout lay{};
  eu::SymPeers peers{};
  char* base = nullptr;
  bool connected = false;
  // local scratch
  int64_t* d_sorted = nullptr; int32_t* d_src = nullptr; int64_t* d_counts = nullptr; int64_t* d_offs = nullptr;
  int64_t* d_rids = nullptr; float* d_rw = nullptr; int32_t* d_rt = nullptr;
  int64_t scratch_rows = 0, scratch_slots = 0;
};

using namespace eu;

extern "C" {

int eu_sym_create(eu_ctx* c, int32_t rank, int32_t world, int64_t max_rows, int32_t max_count, int64_t max_feat_rows,
                  int32_t max_dim, eu_sym** out, void* handle_out /* 64 bytes */) {
  if (!c || !out || !handle_out || world < 1 || world > kSymMaxRanks || rank < 0 || rank >= world || max_rows < 0 ||
      max_count < 1 || max_feat_rows < 0 || max_dim < 0) { set_error("eu_sym_create: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  eu_sym* s = new eu_sym();
  s->c = c; s->rank = rank; s->world = world;
  SymLayout& L = s->lay;
  L.cap = std::max<int64_t>(std::max(max_rows, max_feat_rows), 1);
  L.max_out = max_rows * max_count;
  L.max_rows_f = max_feat_rows;
  L.max_dim = max_dim;
  int64_t off = a256(sizeof(SymHeader));
  L.off_inbox_ids = off; off += a256(8 * L.cap * world);
  L.off_inbox_src = off; off += a256(4 * L.cap * world);
  L.off_eng = off; off += a256(8 * L.max_out);
  L.off_ids = off; off += a256(8 * L.max_out);
  L.off_w = off; off += a256(4 * L.max_out);
  L.off_t = off; off += a256(4 * L.max_out);
  L.off_rows = off; off += a256(4 * max_feat_rows * (int64_t)max_dim);
  L.bytes = off;
  cudaError_t e = cudaMalloc(&s->base, (size_t)L.bytes);
  if (e != cudaSuccess) { set_error("cudaMalloc(%lld) -> %s", (long long)L.bytes, cudaGetErrorString(e)); delete s; return EU_ERR_CUDA; }
  cudaMemset(s->base, 0, (size_t)L.bytes);
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, s->base);
  if (e != cudaSuccess) { set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e)); cudaFree(s->base); delete s; return EU_ERR_CUDA; }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  memcpy(handle_out, &h, 64);
  s->peers.base[rank] = s->base;
  EU_CUDA(cudaDeviceSynchronize());
  *out = s;
  return EU_OK;
}

// handles: world x 64 bytes (all_gather of eu_sym_create's handle_out, rank order)
int eu_sym_connect(eu_sym* s, const void* handles) {
  if (!s || !handles) { set_error("eu_sym_connect: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(s->c->g->device));
  for (int r = 0; r < s->world; ++r) {
    if (r == s->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e)); return EU_ERR_CUDA; }
    s->peers.base[r] = (char*)p;
  }
  s->connected = true;
  return EU_OK;
}

int eu_sym_destroy(eu_sym* s) {
  if (!s) return EU_OK;
  cudaSetDevice(s->c->g->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < s->world; ++r)
    if (r != s->rank && s->peers.base[r]) cudaIpcCloseMemHandle(s->peers.base[r]);
  cudaFree(s->base);
  cudaFree(s->d_sorted); cudaFree(s->d_src); cudaFree(s->d_counts); cudaFree(s->d_offs);
  cudaFree(s->d_rids); cudaFree(s->d_rw); cudaFree(s->d_rt);
  delete s;
  return EU_OK;
}

// device pointers of this rank's output arrays inside the symmetric region (valid until the next call of the same kind)
int eu_sym_outputs(eu_sym* s, int64_t** eng, int64_t** ids, float** w, int32_t** t, float** rows) {
  if (!s) { set_error("null sym"); return EU_ERR_INVALID; }
  if (eng) *eng = (int64_t*)(s->base + s->lay.off_eng);
  if (ids) *ids = (int64_t*)(s->base + s->lay.off_ids);
  if (w) *w = (float*)(s->base + s->lay.off_w);
  if (t) *t = (int32_t*)(s->base + s->lay.off_t);
  if (rows) *rows = (float*)(s->base + s->lay.off_rows);
  return EU_OK;
}

int eu_sym_error(eu_sym* s, int* err) {
  if (!s || !err) { set_error("eu_sym_error: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(s->c->g->device));
  SymHeader h;
  EU_CUDA(cudaMemcpy(&h, s->base, sizeof(h), cudaMemcpyDeviceToHost));
  *err = h.error;
  return EU_OK;
}

static int sym_scratch(eu_sym* s, int64_t rows, int64_t slots) {
  if (rows > s->scratch_rows) {
    EU_CUDA(cudaStreamSynchronize(s->c->stream));
    cudaFree(s->d_sorted); cudaFree(s->d_src);
    EU_CUDA(cudaMalloc(&s->d_sorted, 8 * (size_t)rows));
    EU_CUDA(cudaMalloc(&s->d_src, 4 * (size_t)rows));
    if (!s->d_counts) { EU_CUDA(cudaMalloc(&s->d_counts, 8 * 64)); EU_CUDA(cudaMalloc(&s->d_offs, 8 * 65)); }
    s->scratch_rows = rows;
  }
  if (slots > s->scratch_slots) {
    EU_CUDA(cudaStreamSynchronize(s->c->stream));
    cudaFree(s->d_rids); cudaFree(s->d_rw); cudaFree(s->d_rt);
    EU_CUDA(cudaMalloc(&s->d_rids, 8 * (size_t)slots));
    EU_CUDA(cudaMalloc(&s->d_rw, 4 * (size_t)slots));
    EU_CUDA(cudaMalloc(&s->d_rt, 4 * (size_t)slots));
    s->scratch_slots = slots;
  }
  return EU_OK;
}

// One sharded sampleNB hop.  seeds: this rank's frontier (device, i64[rows]).  Results land in this rank's symmetric
// output arrays (eu_sym_outputs): eng ids (the next frontier) always, TF-packed ids/w/t when want_packed.
int eu_sym_sample_hop(eu_sym* s, const int64_t* seeds, int64_t rows, const int32_t* etypes, int32_t K, int32_t count,
                      int64_t default_node, int32_t num_partitions, int32_t want_packed) {
  if (!s || !s->connected || rows < 0 || count < 0 || (rows > 0 && !seeds)) { set_error("eu_sym_sample_hop: bad argument / not connected"); return EU_ERR_INVALID; }
  eu_ctx* c = s->c;
  EU_CUDA(cudaSetDevice(c->g->device));
  // segment stride of THIS exchange = the requester's row count (every rank issues the same exchange with the same
  // `rows`: batches are equal-sized across ranks), so the padded inbox is N*rows, not N*capacity
  SymLayout L = s->lay;
  const int N = s->world;
  if (rows > L.cap || rows * count > L.max_out) { set_error("eu_sym_sample_hop: %lld rows x %d exceed the symmetric region", (long long)rows, count); return EU_ERR_INVALID; }
  L.cap = std::max<int64_t>(rows, 1);
  const int64_t prow = (int64_t)N * L.cap;   // padded inbox rows
  int rc = sym_scratch(s, std::max<int64_t>(rows, 1), std::max<int64_t>(prow * count, 1));
  if (rc) return rc;
  cudaStream_t st = c->stream;
  rc = eu_shard_bucket(c, seeds, rows, num_partitions, N, s->rank, s->d_sorted, s->d_src, s->d_counts, s->d_offs);
  if (rc) return rc;
  { EuProfScope ps(c, "k_sym_push", rows);
    k_sym_push<<<sym_grid(rows), 256, 0, st>>>(s->peers, L, s->rank, N, (const unsigned long long*)s->d_sorted,
                                                                                    s->d_src, (const long long*)s->d_offs, rows); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_wait_in", rows); k_sym_wait_in<<<1, 32, 0, st>>>(s->base, N); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_wait_pad", prow); k_sym_wait_pad<<<148, 256, 0, st>>>(s->base, L, N); }
  EU_LAUNCHED();
  if (count > 0) {
    rc = hop(c, (const unsigned long long*)(s->base + L.off_inbox_ids), prow, etypes, K, count, /*default_node=*/0, nullptr,
             s->d_rids, s->d_rw, s->d_rt, 0, false, false, 1);
    if (rc) return rc;
  }
  { EuProfScope ps(c, "k_sym_reply_sample", prow);
    k_sym_reply_sample<<<sym_grid(prow * count), 256, 0, st>>>(s->peers, L, s->rank, N, count, default_node,
                                                                                                   (const long long*)s->d_rids, s->d_rw, s->d_rt,
                                                                                                   want_packed != 0); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_wait", rows); k_sym_wait<<<1, 32, 0, st>>>(s->base, N); }
  EU_LAUNCHED();
  return EU_OK;
}

// Sharded dense feature fetch: rows land in this rank's symmetric `rows` output, [rows, dim], request order.
int eu_sym_get_dense_feature(eu_sym* s, const int64_t* ids, int64_t rows, int32_t fid, int32_t dim, int32_t num_partitions) {
  if (!s || !s->connected || rows < 0 || dim <= 0 || (rows > 0 && !ids)) { set_error("eu_sym_get_dense_feature: bad argument / not connected"); return EU_ERR_INVALID; }
  eu_ctx* c = s->c;
  EU_CUDA(cudaSetDevice(c->g->device));
  SymLayout L = s->lay;
  const DevGraph& d = c->g->d;
  const int N = s->world;
  if (rows > L.cap || rows > L.max_rows_f || dim > L.max_dim) { set_error("eu_sym_get_dense_feature: request exceeds the symmetric region"); return EU_ERR_INVALID; }
  const bool have = fid >= 0 && fid < d.n_slots;
  const int32_t soff = have ? d.slot_off[fid] : 0, sdim = have ? d.slot_dim[fid] : 0;
  if ((dim & 3) || (soff & 3) || (sdim & 3) || (d.feat_dim & 3)) { set_error("eu_sym_get_dense_feature: widths must be multiples of 4 floats"); return EU_ERR_UNSUPPORTED; }
  L.cap = std::max<int64_t>(rows, 1);
  int rc = sym_scratch(s, std::max<int64_t>(rows, 1), 1);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  rc = eu_shard_bucket(c, ids, rows, num_partitions, N, s->rank, s->d_sorted, s->d_src, s->d_counts, s->d_offs);
  if (rc) return rc;
  { EuProfScope ps(c, "k_sym_push(feat)", rows);
    k_sym_push<<<sym_grid(rows), 256, 0, st>>>(s->peers, L, s->rank, N, (const unsigned long long*)s->d_sorted,
                                                                                    s->d_src, (const long long*)s->d_offs, rows); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_wait_in(feat)", rows); k_sym_wait_in<<<1, 32, 0, st>>>(s->base, N); }
  EU_LAUNCHED();
  int G = 1;
  while (G < 32 && G < dim / 4) G <<= 1;
  const int64_t prow = (int64_t)N * L.cap;
  { EuProfScope ps(c, "k_sym_reply_feature", prow);
    k_sym_reply_feature<<<sym_grid(prow * G), 256, 0, st>>>(d, s->peers, L, s->rank, N, dim, soff, sdim, G); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_wait(feat)", rows); k_sym_wait<<<1, 32, 0, st>>>(s->base, N); }
  EU_LAUNCHED();
  return EU_OK;
}

// Sharded fused SAGE mean over a fixed-fanout block: out[r,:] = mean_j feat(nbr_ids[r*count+j]) with features fetched
// from the owning shards and summed THERE (see k_sym_reply_sage).  Every rank calls it with the same rows/count/dim.
// Uses the symmetric `rows` region as N partial blocks (clobbers eu_sym_get_dense_feature's output).
int eu_sym_sage_mean(eu_sym* s, const int64_t* nbr_ids, int64_t rows, int32_t count, int32_t dim, int32_t num_partitions, float* out) {
  if (!s || !s->connected || rows < 0 || count < 1 || dim <= 0 || (rows > 0 && (!nbr_ids || !out))) { set_error("eu_sym_sage_mean: bad argument / not connected"); return EU_ERR_INVALID; }
  eu_ctx* c = s->c;
  EU_CUDA(cudaSetDevice(c->g->device));
  SymLayout L = s->lay;
  const DevGraph& d = c->g->d;
  const int N = s->world;
  const int64_t nid = rows * count;
  if (nid > L.cap || nid >= ((int64_t)1 << 31) || (int64_t)N * rows * dim > L.max_rows_f * (int64_t)L.max_dim) {
    set_error("eu_sym_sage_mean: %lld x %d ids / %d partial blocks exceed the symmetric region", (long long)rows, count, N);
    return EU_ERR_INVALID;
  }
  L.cap = std::max<int64_t>(nid, 1);
  int rc = sym_scratch(s, std::max<int64_t>(nid, 1), 1);
  if (rc) return rc;
  cudaStream_t st = c->stream;
  rc = eu_shard_bucket(c, nbr_ids, nid, num_partitions, N, s->rank, s->d_sorted, s->d_src, s->d_counts, s->d_offs);
  if (rc) return rc;
  { EuProfScope ps(c, "k_sym_push(sage)", nid);
    k_sym_push<<<sym_grid(nid), 256, 0, st>>>(s->peers, L, s->rank, N, (const unsigned long long*)s->d_sorted, s->d_src,
                                               (const long long*)s->d_offs, nid); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_wait_in(sage)", nid); k_sym_wait_in<<<1, 32, 0, st>>>(s->base, N); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_reply_sage", (int64_t)N * rows);
    const unsigned grid = sym_grid((int64_t)N * rows * 32);
    const bool fast = d.n < ((int64_t)1 << 31) && d.n_slots == 1 && dim == d.feat_dim;
    if (fast && dim == 128) k_sym_reply_sage<1><<<grid, 256, 0, st>>>(d, s->peers, L, s->rank, N, rows, count);
    else if (fast && dim == 256) k_sym_reply_sage<2><<<grid, 256, 0, st>>>(d, s->peers, L, s->rank, N, rows, count);
    else k_sym_reply_sage_generic<<<grid, 256, 0, st>>>(d, s->peers, L, s->rank, N, rows, count, dim); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_sym_wait(sage)", rows); k_sym_wait<<<1, 32, 0, st>>>(s->base, N); }
  EU_LAUNCHED();
  if (rows > 0) {
    EuProfScope ps(c, "k_sym_sage_reduce", rows);
    k_sym_sage_reduce<<<sym_grid(rows * dim), 256, 0, st>>>((const float*)(s->base + L.off_rows), N, rows, dim, count, out);
    EU_LAUNCHED();
  }
  return EU_OK;
}

}  // extern "C"
