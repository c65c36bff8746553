// Sharding primitives for the multi-GPU path: route seeds to the rank that owns them and merge the
// replies back, replacing the reference's ID_SPLIT -> REMOTE (gRPC) -> IDX_MERGE / DATA_MERGE chain
// (euler/core/kernels/id_split_op.cc:46-99, remote_op.cc:60-146, idx_merge_op.cc:32-78) with two small
// kernels either side of an NCCL all-to-all over NVLink.
//   owner(id) = (id % num_partitions) % shard_num                       (id_split_op.cc:46-49)
// eu_shard_bucket is a STABLE counting sort by owner (the reference keeps batch order inside each shard's
// request, id_split_op.cc:70-75), so every shard sees a deterministic seed order.
#include "internal.h"

namespace eu {

static constexpr int kMaxShards = 64;
static constexpr int kBktBlock = 1024;

// ids 0 (the engine's "no neighbor" placeholder, DEFAULT_UINT64) and 2^64-1 (default_node = -1 fed back as a
// seed) exist on no shard: they resolve to empty rows wherever they are looked up, so they stay on the
// requesting rank instead of all piling onto shard 0 / shard (2^64-1) % N.
__device__ __forceinline__ int owner_of(unsigned long long id, int P, int N, int self) {
  if (id == 0ull || id == ~0ull) return self;
  return (int)((id % (unsigned long long)P) % (unsigned long long)N);
}

// pass 1: per-block histogram; the last block turns blkcnt[b][o] into exclusive bases in (owner, block) order
__global__ void __launch_bounds__(kBktBlock) k_bucket_count(const unsigned long long* __restrict__ ids, int64_t rows, int P, int N, int self,
                                                            uint32_t* blkcnt /*[nblk][N]*/, long long* counts /*[N]*/,
                                                            long long* offsets /*[N+1]*/, unsigned int* done) {
  __shared__ uint32_t s_cnt[kMaxShards];
  __shared__ bool s_last;
  if (threadIdx.x < N) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = blockIdx.x * (int64_t)kBktBlock + threadIdx.x;
  // warp-aggregated: N is small, so a plain shared atomic per row would serialise 256 threads on a few counters
  const int o = i < rows ? owner_of(ids[i], P, N, self) : -1;
  const unsigned same = __match_any_sync(0xffffffffu, o);
  if (o >= 0 && (threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&s_cnt[o], (uint32_t)__popc(same));
  __syncthreads();
  if (threadIdx.x < N) blkcnt[(int64_t)blockIdx.x * N + threadIdx.x] = s_cnt[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // serial over owners (<= 64), parallel over blocks with a running carry: base[b][o]
  __shared__ uint32_t s_scan[kBktBlock];
  uint32_t carry = 0;
  for (int o = 0; o < N; ++o) {
    const uint32_t start = carry;
    for (uint32_t b0 = 0; b0 < gridDim.x; b0 += kBktBlock) {
      const uint32_t b = b0 + threadIdx.x;
      const uint32_t v = b < gridDim.x ? __ldcg(blkcnt + (int64_t)b * N + o) : 0u;
      s_scan[threadIdx.x] = v;
      __syncthreads();
      for (int off = 1; off < kBktBlock; off <<= 1) {
        uint32_t t = threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0u;
        __syncthreads();
        s_scan[threadIdx.x] += t;
        __syncthreads();
      }
      if (b < gridDim.x) blkcnt[(int64_t)b * N + o] = carry + s_scan[threadIdx.x] - v;
      carry += s_scan[kBktBlock - 1];
      __syncthreads();
    }
    if (threadIdx.x == 0) { counts[o] = (long long)(carry - start); offsets[o] = (long long)start; }
  }
  if (threadIdx.x == 0) { offsets[N] = (long long)carry; *done = 0; }
}

// pass 2: stable placement.  rank inside the block = number of earlier lanes / warps with the same owner.
__global__ void __launch_bounds__(kBktBlock) k_bucket_place(const unsigned long long* __restrict__ ids, int64_t rows, int P, int N, int self,
                                                            const uint32_t* __restrict__ base /*[nblk][N]*/,
                                                            unsigned long long* __restrict__ sorted_ids, int32_t* __restrict__ src_index) {
  __shared__ uint32_t s_w[kBktBlock / 32][kMaxShards];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t i = blockIdx.x * (int64_t)kBktBlock + threadIdx.x;
  const bool valid = i < rows;
  const unsigned long long id = valid ? ids[i] : 0ull;
  const int o = valid ? owner_of(id, P, N, self) : -1;
  const unsigned peers = __match_any_sync(0xffffffffu, o);
  const int before = __popc(peers & ((1u << lane) - 1u));
  for (int k = lane; k < N; k += 32) s_w[wid][k] = 0;
  __syncwarp();
  if (valid && before == 0) s_w[wid][o] = __popc(peers);
  __syncthreads();
  if (valid) {
    uint32_t pos = base[(int64_t)blockIdx.x * N + o] + before;
    for (int w = 0; w < wid; ++w) pos += s_w[w][o];
    sorted_ids[pos] = id;
    src_index[pos] = (int32_t)i;
  }
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
  if (rows >= ((int64_t)1 << 31)) { set_error("rows >= 2^31"); return EU_ERR_UNSUPPORTED; }
  const int64_t nblk = rows > 0 ? ceil_div(rows, kBktBlock) : 1;
  int rc = ctx_misc(c, 256 + 4 * nblk * shard_num);
  if (rc) return rc;
  unsigned int* done = (unsigned int*)((char*)c->d_misc + 64);
  uint32_t* blkcnt = (uint32_t*)((char*)c->d_misc + 256);
  cudaStream_t s = c->stream;
  EU_CUDA(cudaMemsetAsync(done, 0, sizeof(unsigned int), s));
  EuProfScope ps(c, "k_bucket(count+place)", rows);
  k_bucket_count<<<(unsigned)nblk, kBktBlock, 0, s>>>((const unsigned long long*)ids, rows, num_partitions, shard_num, self_shard, blkcnt,
                                                      (long long*)counts, (long long*)offsets, done);
  EU_LAUNCHED();
  if (rows > 0) {
    k_bucket_place<<<(unsigned)nblk, kBktBlock, 0, s>>>((const unsigned long long*)ids, rows, num_partitions, shard_num, self_shard, blkcnt,
                                                        (unsigned long long*)sorted_ids, src_index);
    EU_LAUNCHED();
  }
  return EU_OK;
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
