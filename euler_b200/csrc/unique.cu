// Device-side first-occurrence unique: the tf.unique of UniqueDataFlow / SageDataFlow
// (tf_euler/python/dataflow/neighbor_dataflow.py:84-109, sage_dataflow.py:35-50) without leaving HBM.
//   tf.unique(x) -> (y, idx): y = the distinct values of x in order of FIRST occurrence, x[i] == y[idx[i]].
// Same machinery as the seed dedup of the sampling path (csrc/sample.cu): an open-addressing table keyed by id that
// keeps the minimum index, then "is this my first occurrence" flags, an exclusive scan, and a scatter.
//   k_uq_insert -> k_uq_first -> cub::ExclusiveSum -> k_uq_emit        (4 launches, no host sync)
#include <cub/device/device_scan.cuh>

#include "internal.h"

namespace eu {

// slot = {id + 1, min index}; key 0 = free; id 2^64-1 (tag overflow) lives in the extra slot [mask + 1]
__device__ __forceinline__ void uq_insert(HashSlot* tab, unsigned long long mask, unsigned long long id, unsigned long long i) {
  const unsigned long long tag = id + 1;
  if (tag == 0ull) { atomicMin(&tab[mask + 1].row, i); return; }
  unsigned long long h = mix64(id) & mask;
  while (true) {
    const unsigned long long prev = atomicCAS(&tab[h].key, 0ull, tag);
    if (prev == 0ull || prev == tag) { atomicMin(&tab[h].row, i); return; }
    h = (h + 1) & mask;
  }
}

__device__ __forceinline__ unsigned long long uq_first(const HashSlot* tab, unsigned long long mask, unsigned long long id) {
  const unsigned long long tag = id + 1;
  if (tag == 0ull) return tab[mask + 1].row;
  unsigned long long h = mix64(id) & mask;
  while (true) {
    const ulonglong2 s = *reinterpret_cast<const ulonglong2*>(tab + h);
    if (s.x == tag) return s.y;
    h = (h + 1) & mask;
  }
}

__global__ void k_uq_clear(HashSlot* tab, int64_t slots) {
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < slots; s += (int64_t)gridDim.x * blockDim.x) {
    tab[s].key = 0ull; tab[s].row = kEmptyRow;
  }
}

__global__ void k_uq_insert(HashSlot* tab, unsigned long long mask, const unsigned long long* __restrict__ ids, int64_t n) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long id = ids[i];
  // runs of equal ids (default fill, hubs) would hammer one slot: the lowest lane of each group carries the group's minimum
  const unsigned peers = __match_any_sync(__activemask(), id);
  if ((threadIdx.x & 31) != __ffs(peers) - 1) return;
  uq_insert(tab, mask, id, (unsigned long long)i);
}

__global__ void k_uq_first(const HashSlot* tab, unsigned long long mask, const unsigned long long* __restrict__ ids, int64_t n,
                           int32_t* __restrict__ first, int32_t* __restrict__ flag) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t f = (int32_t)uq_first(tab, mask, ids[i]);
  first[i] = f;
  flag[i] = f == (int32_t)i ? 1 : 0;
}

__global__ void k_uq_emit(const unsigned long long* __restrict__ ids, int64_t n, const int32_t* __restrict__ first,
                          const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, unsigned long long* __restrict__ uniq,
                          int32_t* __restrict__ inverse, long long* __restrict__ n_unique) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (flag[i]) uniq[pos[i]] = ids[i];
  inverse[i] = pos[first[i]];
  if (i == n - 1 && n_unique) *n_unique = (long long)pos[i] + flag[i];
}

}  // namespace eu

using namespace eu;

// uniq: device i64[n] (first *n_unique entries valid), inverse: device i32[n], n_unique: device i64[1] (may be NULL)
extern "C" int eu_unique(eu_ctx* c, const int64_t* ids, int64_t n, int64_t* uniq, int32_t* inverse, int64_t* n_unique) {
  if (!c || n < 0 || n >= ((int64_t)1 << 31) || (n > 0 && (!ids || !uniq || !inverse))) { set_error("eu_unique: bad argument"); return EU_ERR_INVALID; }
  EU_CUDA(cudaSetDevice(c->g->device));
  cudaStream_t s = c->stream;
  if (n == 0) {
    if (n_unique) EU_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(int64_t), s));
    return EU_OK;
  }
  int64_t cap = 64;
  while (cap < 2 * n) cap <<= 1;
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum((void*)nullptr, tmp, (int32_t*)nullptr, (int32_t*)nullptr, (int)n, s);
  const int64_t o_tab = 256, o_first = o_tab + 16 * (cap + 1), o_flag = o_first + ((4 * n + 255) & ~(int64_t)255),
                o_pos = o_flag + ((4 * n + 255) & ~(int64_t)255), o_tmp = o_pos + ((4 * n + 255) & ~(int64_t)255);
  int rc = ctx_misc(c, o_tmp + (int64_t)tmp + 256);
  if (rc) return rc;
  char* m = (char*)c->d_misc;
  HashSlot* tab = (HashSlot*)(m + o_tab);
  int32_t *first = (int32_t*)(m + o_first), *flag = (int32_t*)(m + o_flag), *pos = (int32_t*)(m + o_pos);
  const unsigned nb = (unsigned)ceil_div(n, 256);
  { EuProfScope ps(c, "k_uq_clear+insert", n);
    k_uq_clear<<<(unsigned)std::min<int64_t>(ceil_div(cap + 1, 256), 148 * 8), 256, 0, s>>>(tab, cap + 1);
    EU_LAUNCHED();
    k_uq_insert<<<nb, 256, 0, s>>>(tab, (unsigned long long)cap - 1, (const unsigned long long*)ids, n); }
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_uq_first", n);
    k_uq_first<<<nb, 256, 0, s>>>(tab, (unsigned long long)cap - 1, (const unsigned long long*)ids, n, first, flag); }
  EU_LAUNCHED();
  EU_CUDA(cub::DeviceScan::ExclusiveSum(m + o_tmp, tmp, flag, pos, (int)n, s));
  EU_LAUNCHED();
  { EuProfScope ps(c, "k_uq_emit", n);
    k_uq_emit<<<nb, 256, 0, s>>>((const unsigned long long*)ids, n, first, flag, pos, (unsigned long long*)uniq, inverse, (long long*)n_unique); }
  EU_LAUNCHED();
  return EU_OK;
}
