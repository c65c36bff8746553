// Sharding primitives for the multi-GPU path: route seeds to the rank that owns them and merge the
// replies back, replacing the reference's ID_SPLIT -> REMOTE (gRPC) -> IDX_MERGE / DATA_MERGE chain
// (euler/core/kernels/id_split_op.cc:46-99, remote_op.cc:60-146, idx_merge_op.cc:32-78) with two small
// kernels either side of an NCCL all-to-all over NVLink.
//   owner(id) = (id % num_partitions) % shard_num                       (id_split_op.cc:46-49)
// eu_shard_bucket is a STABLE counting sort by owner (the reference keeps batch order inside each shard's
// request, id_split_op.cc:70-75), so every shard sees a deterministic seed order.
#include "internal.h"
#include "sym.cuh"

namespace eu {

static constexpr int kMaxShards = 64;
static constexpr int kBktWarps = 8;     // warps per CTA
static constexpr int kChunk = 256;      // ids per warp: 8 coalesced rounds, all loads in flight before the first use

// ids 0 (the engine's "no neighbor" placeholder, DEFAULT_UINT64) and 2^64-1 (default_node = -1 fed back as a
// seed) exist on no shard: they resolve to empty rows wherever they are looked up, so they stay on the
// requesting rank instead of all piling onto shard 0 / shard (2^64-1) % N -- or are dropped (-1) when the caller
// needs nothing back for them (the fused aggregation).
__device__ __forceinline__ int owner_of(unsigned long long id, int P, int N, int self, bool drop) {
  if (id == 0ull || id == ~0ull) return drop ? -1 : self;
  return (int)((id % (unsigned long long)P) % (unsigned long long)N);
}

// Stable counting sort by owner.  Warp w of CTA b owns ids [(b*8+w)*256, +256).
// pass 1: per-CTA histogram bcnt[b][o]; the last CTA turns it into each CTA's exclusive base INSIDE its owner's segment
// (strip-serial + one block scan per owner) and writes counts[o] / offsets[o] (segment starts in the sorted array).
__global__ void __launch_bounds__(kBktWarps * 32) k_bucket_count(const unsigned long long* __restrict__ ids, int64_t rows, int P, int N,
                                                                 int self, bool drop, uint32_t* bcnt /*[nblk][N]*/,
                                                                 long long* counts /*[N]*/, long long* offsets /*[N+1]*/,
                                                                 unsigned int* done) {
  __shared__ uint32_t s_cnt[kMaxShards];
  __shared__ uint32_t s_scan[kBktWarps * 32];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < N) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  {
    unsigned long long v[kChunk / 32];
    const int64_t i0 = (blockIdx.x * (int64_t)kBktWarps + wid) * kChunk + lane;
#pragma unroll
    for (int r = 0; r < kChunk / 32; ++r) v[r] = i0 + r * 32 < rows ? ids[i0 + r * 32] : 0ull;
#pragma unroll
    for (int r = 0; r < kChunk / 32; ++r) {
      const int o = i0 + r * 32 < rows ? owner_of(v[r], P, N, self, drop) : -1;
      const unsigned m = __match_any_sync(0xffffffffu, o);
      if (o >= 0 && lane == __ffs(m) - 1) atomicAdd(&s_cnt[o], (uint32_t)__popc(m));   // one leader per (warp, owner)
    }
  }
  __syncthreads();
  if (threadIdx.x < N) bcnt[(int64_t)blockIdx.x * N + threadIdx.x] = s_cnt[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); s_last = atomicAdd(done, 1u) == gridDim.x - 1; }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int64_t nblk = gridDim.x;
  const int64_t strip = (nblk + blockDim.x - 1) / blockDim.x;
  const int64_t b = min((int64_t)threadIdx.x * strip, nblk), e = min(b + strip, nblk);
  uint32_t start = 0;
  for (int o = 0; o < N; ++o) {
    uint32_t sum = 0;
    for (int64_t k = b; k < e; ++k) sum += __ldcg(bcnt + k * N + o);
    s_scan[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < kBktWarps * 32; off <<= 1) {
      const uint32_t t = (int)threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0u;
      __syncthreads();
      s_scan[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t run = s_scan[threadIdx.x] - sum;   // exclusive base of this thread's strip inside owner o's segment
    const uint32_t total = s_scan[kBktWarps * 32 - 1];
    for (int64_t k = b; k < e; ++k) {
      const uint32_t c = __ldcg(bcnt + k * N + o);
      bcnt[k * N + o] = run;
      run += c;
    }
    if (threadIdx.x == 0) { counts[o] = (long long)total; offsets[o] = (long long)start; }
    start += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) { offsets[N] = (long long)start; *done = 0; }
}

// pass 2: every id goes to its owner's segment at CTA base + (ids of the same owner before it in the CTA).
// Local mode: segment o = sorted[offsets[o] ..).  Remote mode (the peer-memory exchange): segment o IS owner o's inbox
// slice for this rank, written over NVLink, and the last CTA publishes the counts and raises flagA -- the bucket and
// the push are one kernel.
struct BucketDst {
  unsigned long long* sorted_ids;   // local mode
  int32_t* src_index;
  int remote;
  char* const* pb_tab;              // remote mode: device table of the peers' region bases
  SymLayout lay;
  int me;
};

__global__ void __launch_bounds__(kBktWarps * 32) k_bucket_place(const unsigned long long* __restrict__ ids, int64_t rows, int P, int N,
                                                                 int self, bool drop, const uint32_t* __restrict__ bbase,
                                                                 const long long* __restrict__ counts,
                                                                 const long long* __restrict__ offsets, BucketDst dst) {
  __shared__ uint32_t s_w[kBktWarps][kMaxShards];   // per-warp counts, then per-warp running bases
  __shared__ bool s_last;
  __shared__ int s_err;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (dst.remote) {   // a timed-out exchange poisons the region (p2p.cu): push nothing, raise nothing
    if (threadIdx.x == 0) s_err = ld_volatile_i32(&hdr_of(dst.pb_tab[dst.me])->error);
    __syncthreads();
    if (s_err) return;
  }
  for (int k = lane; k < N; k += 32) s_w[wid][k] = 0;
  __syncwarp();
  unsigned long long v[kChunk / 32];
  int8_t ow[kChunk / 32];
  unsigned mm[kChunk / 32];
  const int64_t i0 = (blockIdx.x * (int64_t)kBktWarps + wid) * kChunk + lane;
#pragma unroll
  for (int r = 0; r < kChunk / 32; ++r) v[r] = i0 + r * 32 < rows ? ids[i0 + r * 32] : 0ull;
#pragma unroll
  for (int r = 0; r < kChunk / 32; ++r) {
    const int o = i0 + r * 32 < rows ? owner_of(v[r], P, N, self, drop) : -1;
    const unsigned m = __match_any_sync(0xffffffffu, o);
    ow[r] = (int8_t)o; mm[r] = m;
    if (o >= 0 && lane == __ffs(m) - 1) s_w[wid][o] += (uint32_t)__popc(m);
    __syncwarp();
  }
  __syncthreads();
  // running base of warp wid for owner k = CTA base + counts of the warps before it
  if (threadIdx.x < N) {
    uint32_t run = bbase[(int64_t)blockIdx.x * N + threadIdx.x];
    for (int w = 0; w < kBktWarps; ++w) { const uint32_t c = s_w[w][threadIdx.x]; s_w[w][threadIdx.x] = run; run += c; }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kChunk / 32; ++r) {
    const int o = ow[r];
    const unsigned m = mm[r];
    uint32_t rel = 0;
    if (o >= 0) rel = s_w[wid][o] + (uint32_t)__popc(m & ((1u << lane) - 1u));
    __syncwarp();
    if (o >= 0 && lane == __ffs(m) - 1) s_w[wid][o] += (uint32_t)__popc(m);
    __syncwarp();
    if (o >= 0) {
      const int64_t i = i0 + r * 32;
      if (dst.remote) {
        char* pb = dst.pb_tab[o];
        reinterpret_cast<unsigned long long*>(pb + dst.lay.off_inbox_ids)[(int64_t)dst.me * dst.lay.cap + rel] = v[r];
        reinterpret_cast<int32_t*>(pb + dst.lay.off_inbox_src)[(int64_t)dst.me * dst.lay.cap + rel] = (int32_t)i;
      } else {
        const int64_t pos = offsets[o] + rel;
        dst.sorted_ids[pos] = v[r];
        dst.src_index[pos] = (int32_t)i;
      }
    }
  }
  if (!dst.remote) return;
  // publish: one system fence per CTA (it waits for the NVLink acks of the CTA's stores), last CTA raises the flags
  __syncthreads();   // CTA's stores happen-before thread 0's fence (barrier + cumulativity)
  SymHeader* mine = hdr_of(dst.pb_tab[dst.me]);
  if (threadIdx.x == 0) { __threadfence_system(); s_last = atomicAdd(&mine->done, 1u) == gridDim.x - 1; }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x < N) __threadfence_system();   // ticket observed (barrier) -> ordered before the flag stores
  __shared__ unsigned int s_e;
  if (threadIdx.x == 0) { s_e = mine->epoch + 1; mine->epoch = s_e; mine->done = 0; }
  __syncthreads();
  if (threadIdx.x < N) {
    SymHeader* h = hdr_of(dst.pb_tab[threadIdx.x]);
    h->in_cnt[dst.me] = (int)counts[threadIdx.x];
    h->in_total[dst.me] = (int)rows;
    __threadfence_system();
    st_release_sys(&h->flagA[dst.me], s_e);
  }
}

static int bucket_launch(eu_ctx* c, const int64_t* ids, int64_t rows, int P, int N, int self, bool drop, int64_t* counts,
                         int64_t* offsets, const BucketDst& dst, const char* label) {
  if (rows >= ((int64_t)1 << 31)) { set_error("rows >= 2^31"); return EU_ERR_UNSUPPORTED; }
  const int64_t nblk = rows > 0 ? ceil_div(rows, (int64_t)kChunk * kBktWarps) : 1;
  int rc = ctx_misc(c, 256 + 4 * nblk * N);
  if (rc) return rc;
  unsigned int* done = (unsigned int*)((char*)c->d_misc + 64);
  uint32_t* wcnt = (uint32_t*)((char*)c->d_misc + 256);
  cudaStream_t s = c->stream;
  EU_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned int), s));
  EuProfScope ps(c, label, rows);
  k_bucket_count<<<(unsigned)nblk, kBktWarps * 32, 0, s>>>((const unsigned long long*)ids, rows, P, N, self, drop, wcnt,
                                                           (long long*)counts, (long long*)offsets, done);
  EU_LAUNCHED();
  if (rows > 0 || dst.remote) {   // remote mode always runs: the flags must be raised even for an empty request
    k_bucket_place<<<(unsigned)nblk, kBktWarps * 32, 0, s>>>((const unsigned long long*)ids, rows, P, N, self, drop, wcnt,
                                                             (const long long*)counts, (const long long*)offsets, dst);
    EU_LAUNCHED();
  }
  return EU_OK;
}

// bucket + push of the peer-memory exchange (p2p.cu): requests land in the owners' inboxes, counts and flagA follow
int bucket_push(eu_ctx* c, const int64_t* ids, int64_t rows, int P, int N, int self, bool drop_placeholders, int64_t* counts,
                int64_t* offsets, char* const* pb_tab, const SymLayout& lay, const char* label) {
  BucketDst dst{};
  dst.remote = 1; dst.pb_tab = pb_tab; dst.lay = lay; dst.me = self;
  return bucket_launch(c, ids, rows, P, N, self, drop_placeholders, counts, offsets, dst, label);
}

// reply merge + TF packing for sampled rows: reply row k (sorted order) belongs to original row src_index[k].
// eng ids (0 = placeholder) go to the next frontier; packed outputs get default_node / 0 / -1 when the row's
// first id is 0 (tf_euler/kernels/sample_neighbor_op.cc:79-81,114-122).
// SoA (ids, w, t) -> one 16-byte record per slot {id, w bits | t << 32}: a single all-to-all carries the reply
__global__ void k_pack_rows(const long long* __restrict__ ids, const float* __restrict__ w, const int32_t* __restrict__ t,
                            int64_t n, longlong2* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  longlong2 v;
  v.x = ids[i];
  v.y = (long long)(((unsigned long long)(uint32_t)t[i] << 32) | (unsigned long long)__float_as_uint(w[i]));
  out[i] = v;
}

__global__ void k_merge_sample(const longlong2* __restrict__ rec,
                               const int32_t* __restrict__ src_index, int64_t rows, int32_t count, long long default_node,
                               unsigned long long* __restrict__ eng_ids, long long* __restrict__ out_ids,
                               float* __restrict__ out_w, int32_t* __restrict__ out_t) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tid >= rows * count) return;
  const int64_t k = tid / count;
  const int32_t j = (int32_t)(tid % count);
  const int64_t dst = (int64_t)src_index[k] * count + j;
  const longlong2 v = rec[tid];
  const long long id = v.x;
  const bool keep = rec[k * count].x != 0;
  if (eng_ids) eng_ids[dst] = (unsigned long long)id;
  if (out_ids) {
    out_ids[dst] = keep ? id : default_node;
    out_w[dst] = keep ? __uint_as_float((uint32_t)((unsigned long long)v.y & 0xffffffffull)) : 0.f;
    out_t[dst] = keep ? (int32_t)((unsigned long long)v.y >> 32) : -1;
  }
}

// reply merge for fixed-width f32 rows (features): out[src_index[k], :] = rows[k, :]
__global__ void k_merge_rows(const float* __restrict__ in, const int32_t* __restrict__ src_index, int64_t rows, int64_t D,
                             int G, float* __restrict__ out) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t k = tid >> (31 - __clz(G));   // G is a power of two
  const int sub = (int)(tid & (G - 1));
  if (k >= rows) return;
  const float* s = in + k * D;
  float* o = out + (int64_t)src_index[k] * D;
  if ((D & 3) == 0) {
    for (int64_t d = sub * 4; d < D; d += G * 4) *reinterpret_cast<float4*>(o + d) = __ldg(reinterpret_cast<const float4*>(s + d));
  } else {
    for (int64_t d = sub; d < D; d += G) o[d] = __ldg(s + d);
  }
}

}  // namespace eu

using namespace eu;

extern "C" {

int eu_shard_bucket(eu_ctx* c, const int64_t* ids, int64_t rows, int32_t num_partitions, int32_t shard_num, int32_t self_shard,
                    int64_t* sorted_ids, int32_t* src_index, int64_t* counts, int64_t* offsets) {
  if (!c || rows < 0 || num_partitions <= 0 || shard_num <= 0 || shard_num > kMaxShards || self_shard < 0 || self_shard >= shard_num || !counts || !offsets ||
      (rows > 0 && (!ids || !sorted_ids || !src_index))) {
    set_error("eu_shard_bucket: bad argument");
    return EU_ERR_INVALID;
  }
  EU_CUDA(cudaSetDevice(c->g->device));
  BucketDst dst{};
  dst.sorted_ids = (unsigned long long*)sorted_ids; dst.src_index = src_index;
  return bucket_launch(c, ids, rows, num_partitions, shard_num, self_shard, false, counts, offsets, dst, "k_bucket(count+place)");
}

int eu_shard_pack_sample(eu_ctx* c, const int64_t* ids, const float* w, const int32_t* t, int64_t n, int64_t* packed) {
  if (!c || n < 0 || (n > 0 && (!ids || !w || !t || !packed))) { set_error("eu_shard_pack_sample: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (n == 0) return EU_OK;
  k_pack_rows<<<(unsigned)ceil_div(n, 256), 256, 0, c->stream>>>((const long long*)ids, w, t, n, (longlong2*)packed);
  EU_LAUNCHED();
  return EU_OK;
}

int eu_shard_merge_sample(eu_ctx* c, const int64_t* packed,
                          const int32_t* src_index, int64_t rows, int32_t count, int64_t default_node, int64_t* eng_ids,
                          int64_t* out_ids, float* out_w, int32_t* out_t) {
  if (!c || rows < 0 || count < 0 || (rows * count > 0 && (!packed || !src_index))) { set_error("eu_shard_merge_sample: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (rows * count == 0) return EU_OK;
  k_merge_sample<<<(unsigned)ceil_div(rows * count, 256), 256, 0, c->stream>>>((const longlong2*)packed, src_index, rows, count,
                                                                               default_node, (unsigned long long*)eng_ids, (long long*)out_ids, out_w, out_t);
  EU_LAUNCHED();
  return EU_OK;
}

int eu_shard_merge_rows(eu_ctx* c, const float* rows_in, const int32_t* src_index, int64_t rows, int64_t D, float* out) {
  if (!c || rows < 0 || D <= 0 || (rows > 0 && (!rows_in || !src_index || !out))) { set_error("eu_shard_merge_rows: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (rows == 0) return EU_OK;
  int G = 1;
  if ((D & 3) == 0) { while (G < 32 && G < D / 4) G <<= 1; } else if (D >= 32) G = 32;
  k_merge_rows<<<(unsigned)ceil_div(rows * G, 256), 256, 0, c->stream>>>(rows_in, src_index, rows, D, G, out);
  EU_LAUNCHED();
  return EU_OK;
}

}  // extern "C"
