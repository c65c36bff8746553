// euler_b200 device-side building blocks (sm_100a).
//   * DevGraph: the HBM-resident CSR a kernel sees
//   * exact RNG: minstd_rand0 + libstdc++ generate_canonical<double,53>, with multiplicative
//     jump-ahead so each warp/lane starts at its own position of the ONE serial stream the
//     reference consumes (euler/common/random.cc:22-28)
//   * Philox4x32-10 keyed on (node id, draw) for the throughput mode
//   * inverse-CDF pick equivalent to RandomSelect (euler/common/compact_weighted_collection.h:30-52)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EU_WARP 32
#define EU_MAX_ETYPES 32
#define EU_MAX_FEAT_SLOTS 16

namespace eu {

struct HashSlot {  // 16 B, one 128-bit load
  unsigned long long key;
  unsigned long long row;  // EU_EMPTY_ROW if free
};
static constexpr unsigned long long kEmptyRow = 0xFFFFFFFFFFFFFFFFull;

struct DevGraph {
  int64_t n;        // rows
  int64_t E;        // edges
  int32_t T;        // edge-type groups per row
  int32_t n_node_types;
  const unsigned long long* ids;  // [n]
  const int32_t* node_type;       // [n]
  const float* node_w;            // [n]
  const int64_t* grp_ptr;         // [n*T+1]
  const unsigned long long* nbr;  // [E]
  const float* cum_w;             // [E]
  const float* grp_cum;           // [n*T] (T>1) or nullptr
  // id -> row
  int32_t adj_sorted;             // every adjacency group is non-decreasing in (signed) neighbor id
  int32_t dense_ids;              // ids[r] == id_base + r * id_stride for all r
  unsigned long long id_base;
  unsigned long long id_stride;   // 1, or the shard count for a shard's rows (ids congruent mod shards)
  const HashSlot* htab;
  unsigned long long hmask;       // capacity-1 (capacity is a power of two)
  // dense f32 features: row-major [n, feat_dim]; slot s occupies columns [slot_off[s], +slot_dim[s])
  int32_t feat_dim;
  const float* feat;
  int32_t n_slots;
  int32_t slot_off[EU_MAX_FEAT_SLOTS];
  int32_t slot_dim[EU_MAX_FEAT_SLOTS];
  // ragged features (Node::uint64_features_ / binary_features_ with their *_idx_ ends, node.h): slot s of row r is
  // [ptr[r*S+s], ptr[r*S+s+1]) of the value array; S = 0 when the graph has none
  int32_t n_u64_slots;
  const int64_t* u64_ptr;
  const unsigned long long* u64_val;
  int32_t n_bin_slots;
  const int64_t* bin_ptr;
  const unsigned char* bin_val;
};

// edge store (edges.cu): SoA edge arrays + (src, dst, type) -> row table + features with the node layout
struct EdgeSlot {   // 32 B
  unsigned long long src, dst;
  int32_t type, pad;
  long long row;    // -1 = free
};
struct DevEdges {
  int64_t n;
  const unsigned long long* src;
  const unsigned long long* dst;
  const int32_t* type;
  const float* w;
  const EdgeSlot* htab;
  unsigned long long hmask;
  int32_t feat_dim;
  const float* feat;
  int32_t n_slots;
  int32_t slot_off[EU_MAX_FEAT_SLOTS];
  int32_t slot_dim[EU_MAX_FEAT_SLOTS];
  int32_t n_u64_slots;
  const int64_t* u64_ptr;
  const unsigned long long* u64_val;
  int32_t n_bin_slots;
  const int64_t* bin_ptr;
  const unsigned char* bin_val;
};

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

// Graph::GetNodeByID (euler/core/graph/graph.h:87-93): row or -1.
__device__ __forceinline__ int64_t lookup_row(const DevGraph& g, unsigned long long id) {
  if (g.dense_ids) {
    unsigned long long r = id - g.id_base;
    if (g.id_stride != 1) {
      if (id < g.id_base || r % g.id_stride) return -1;
      r /= g.id_stride;
    }
    return r < (unsigned long long)g.n ? (int64_t)r : -1;
  }
  unsigned long long h = mix64(id) & g.hmask;
  while (true) {
    const ulonglong2 s = __ldg(reinterpret_cast<const ulonglong2*>(g.htab + h));
    if (s.y == kEmptyRow) return -1;
    if (s.x == id) return (int64_t)s.y;
    h = (h + 1) & g.hmask;
  }
}

// ---------------------------------------------------------------------------------- minstd_rand0
static constexpr uint32_t kM = 2147483647u;  // 2^31-1
static constexpr uint32_t kA = 16807u;

__host__ __device__ __forceinline__ uint32_t modmul(uint32_t a, uint32_t b) {
  unsigned long long p = (unsigned long long)a * b;     // < 2^62
  unsigned long long r = (p & kM) + (p >> 31);          // < 2^32
  r = (r & kM) + (r >> 31);
  return (uint32_t)(r >= kM ? r - kM : r);
}

__host__ __device__ __forceinline__ uint32_t modpow_a(unsigned long long e) {  // A^e mod M
  uint32_t base = kA, acc = 1;
  while (e) {
    if (e & 1) acc = modmul(acc, base);
    base = modmul(base, base);
    e >>= 1;
  }
  return acc;
}

// A^(e) for small e (< 256) with compile-time squarings: used for per-lane offsets.
__device__ __forceinline__ uint32_t modpow_a_small(uint32_t e) {
  // A^(2^k) mod M, k = 0..7
  constexpr uint32_t P0 = 16807u;
  uint32_t acc = 1, base = P0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (e & (1u << k)) acc = modmul(acc, base);
    base = modmul(base, base);
  }
  return acc;
}

// One uniform_real_distribution<double>(0,1) draw from engine state `x` (state BEFORE the draw);
// advances x by two engine steps.  Arithmetic order of libstdc++ 13 generate_canonical<double,53>:
//   sum = double(x1-1); sum += double(x2-1) * R; ret = sum / (R*R); ret >= 1 -> nextafter(1,0)
// with R = 2147483646.0.  Explicit _rn intrinsics: no FMA contraction allowed (Appendix A-13/15).
__device__ __forceinline__ double minstd_uniform(uint32_t& x) {
  const double R = 2147483646.0;
  const double RR = 4611686009837453316.0;  // fl(R*R)
  x = modmul(x, kA);
  double sum = (double)(x - 1u);
  x = modmul(x, kA);
  sum = __dadd_rn(sum, __dmul_rn((double)(x - 1u), R));
  double ret = __ddiv_rn(sum, RR);
  if (ret >= 1.0) ret = 0x1.fffffffffffffp-1;  // nextafter(1.0, 0.0)
  return ret;
}

// ---------------------------------------------------------------------------------- per-call seed dedup table
// tab[h] = {id+1, min index}.  key 0 = free.
__device__ __forceinline__ void dedup_insert_one(HashSlot* tab, unsigned long long mask, unsigned long long id,
                                                 int64_t i) {
  const unsigned long long tag = id + 1;
  if (tag == 0ull) {  // id == 2^64-1 (e.g. default_node -1 fed back as a seed): dedicated slot [mask+1]
    atomicMin(&tab[mask + 1].row, (unsigned long long)i);
    return;
  }
  unsigned long long h = mix64(id) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&tab[h].key, 0ull, tag);
    if (prev == 0ull || prev == tag) {
      atomicMin(&tab[h].row, (unsigned long long)i);
      return;
    }
    h = (h + 1) & mask;
  }
}


// Slots a batch's dedup region really uses: the power of two >= 2 * live rows (>= 64), at most the region's capacity.  A
// sharded owner sizes its regions for the worst case (every request of every rank) and learns the live count on the device:
// inserting, probing and wiping with this mask keeps the table -- and its wipe -- proportional to the rows that exist.
__device__ __forceinline__ int64_t dedup_cap_eff(int64_t cap_b, const int32_t* __restrict__ rows_act, int b) {
  if (!rows_act) return cap_b;
  const long long r2 = 2ll * (long long)rows_act[b];
  int64_t cap = 64;
  if (r2 > 64) cap = (int64_t)1 << (64 - __clzll(r2 - 1));
  return cap < cap_b ? cap : cap_b;
}

// ---------------------------------------------------------------------------------- philox4x32-10
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// 2 uniforms in [0,1) with 53 bits from counter (id, draw) and key (seed, call)
__device__ __forceinline__ void philox_uniform2(unsigned long long id, uint32_t draw, uint32_t salt,
                                                unsigned long long key, double& u0, double& u1) {
  uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), draw, salt};
  uint32_t k0 = (uint32_t)key, k1 = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  unsigned long long a = ((unsigned long long)c[0] << 32) | c[1];
  unsigned long long b = ((unsigned long long)c[2] << 32) | c[3];
  u0 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
  u1 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

// ---------------------------------------------------------------------------------- TMA bulk copy (sm_90+/sm_100a)
// 1-D bulk copy global -> shared through the TMA unit, completion signalled on an mbarrier (SASS: UBLKCP + SYNCS).  One
// elected lane issues it and moves on; the consumers wait on the barrier's phase parity.  src, dst and bytes must be
// multiples of 16.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy accesses to shared memory ordered before subsequent async-proxy (TMA) accesses of the same thread
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// ---------------------------------------------------------------------------------- CDF pick
// r for a pick in [begin,end] of a cumulative array (compact_weighted_collection.h:33-36):
//   limit_begin = begin==first ? 0 : cum[begin-1]; r = u * (double)(limit_end - limit_begin) + limit_begin
// (f32 subtraction, then f64 multiply and f64 add, unfused).
__device__ __forceinline__ double pick_r(double u, float limit_begin, float limit_end) {
  float diff = __fsub_rn(limit_end, limit_begin);
  return __dadd_rn(__dmul_rn(u, (double)diff), (double)limit_begin);
}

// Smallest f32 whose value exceeds r (r >= 0: cumulative weights are non-negative).  For an f32 c,
// (double)c > r  <=>  c >= gt_threshold(r), which turns every probe of the search into one FSETP instead of
// F2F.F64 + DSETP.
__device__ __forceinline__ float gt_threshold(double r) {
  float f = __double2float_ru(r);                       // (double)f >= r
  if ((double)f == r) f = __int_as_float(__float_as_int(f) + 1);  // next f32 up (valid for f >= +0)
  return f;
}

// RandomSelect == min(end, first j in [begin,end] with (double)cum[j] > r) for non-decreasing cum
// (proof sketch in DESIGN.md; checked against the literal search in tests/test_oracle_golden.py).
// Search over a row slice in global memory; `base` points at the slice, offsets are 32-bit (a row has
// < 2^31 edges), thr = gt_threshold(r).  Returns the offset in [lo, hi].
// Invariant: the answer lies in [lo, hi]; hi is the clamp or an offset with base[hi] >= thr.
// 8-ary descent: 7 independent probes per round trip.  The rows that land here are hubs (a hop-2 frontier is
// degree-biased) whose slices live in L2; the profile (profiles/r01_*) shows the kernel stalled on exactly this
// load-compare chain, with issue slots to spare for the extra probes.
__device__ __forceinline__ int32_t upper_bound_clamped(const float* __restrict__ base, int32_t lo, int32_t hi,
                                                       float thr) {
  while (hi - lo >= 8) {
    // probes p_i = lo + (i+1)*s + min(i+1, r), s = n/8, r = n%8: lo < p_0 < .. < p_6 < hi (any increasing probe set
    // gives the same answer; this one needs no 64-bit multiply)
    const int32_t n = hi - lo, s = n >> 3, r = n & 7;
    float v[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) v[i] = __ldg(base + lo + (i + 1) * s + min(i + 1, r));
    int c = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) c += (v[i] >= thr) ? 0 : 1;                                          // monotone: a prefix is below thr
    const int32_t nlo = c > 0 ? lo + c * s + min(c, r) + 1 : lo;
    const int32_t nhi = c < 7 ? lo + (c + 1) * s + min(c + 1, r) : hi;
    lo = nlo; hi = nhi;
  }
  if (lo < hi) {  // fewer than 8 candidates below hi: one round of independent probes
    const int32_t n = hi - lo;
    int c = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const float v = i < n ? __ldg(base + lo + i) : __int_as_float(0x7f800000);
      c += (v >= thr) ? 0 : 1;
    }
    lo += c;
  }
  return lo;
}

}  // namespace eu
