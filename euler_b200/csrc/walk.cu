// random_walk: uniform (p=q=1) and node2vec-biased walks, walkers resident on the device for all L steps.
//
// Reference (file:line relative to /root/reference):
//   RandomWalk::ComputeAsync          tf_euler/kernels/random_walk_op.cc:249-289
//   TraditionalRandomWalk (p=q=1)     :207-247  = L chained sampleNB(count=1) hops, id 0 -> default_node
//   RWCallback::operator() (node2vec) :83-138   per step: full neighbor list of the current node
//                                               (Node::__GetFullNeighbor, euler/core/graph/node.cc:176-198),
//   BuildWeights                      :140-168  two-pointer merge against the parent's list,
//   CompactWeightedCollection::Init/Sample      euler/common/compact_weighted_collection.h:82-128
//                                               sequential f32 prefix + one RandomSelect draw.
// The reference makes one host round trip (a GQL query) per step; here a step is two small kernels
// and the frontier never leaves HBM.  Exact-RNG order: one uniform per LIVE walker, in walker order.
#include "internal.h"

namespace eu {

struct ETypes2 {
  int32_t K;
  int32_t v[EU_MAX_ETYPES];
};

// Sequential view of Node::GetFullNeighbor(edge_types): listed types in listed order (invalid ones
// skipped, duplicates repeated), stored order within a group; weight = cum[j] - cum[j-1] (0 at row start).
struct NbIter {
  const DevGraph* g;
  const int64_t* gp;  // grp_ptr + row*T
  int64_t base;
  const ETypes2* et;
  int32_t k;          // position in et
  int64_t j, jend;
  __device__ void init(const DevGraph* g_, int64_t row, const ETypes2* et_) {
    g = g_; et = et_; k = -1; j = 0; jend = 0;
    if (row < 0) { gp = nullptr; k = et_->K; return; }
    gp = g->grp_ptr + row * g->T;
    base = gp[0];
    advance_group();
  }
  __device__ void advance_group() {
    while (j >= jend) {
      ++k;
      if (k >= et->K) return;
      int32_t t = et->v[k];
      if (t < 0 || t >= g->T) continue;
      j = gp[t];
      jend = gp[t + 1];
    }
  }
  __device__ bool done() const { return k >= et->K; }
  __device__ long long id() const { return (long long)__ldg(g->nbr + j); }
  __device__ float w() const {
    float hi = __ldg(g->cum_w + j);
    float lo = j == base ? 0.f : __ldg(g->cum_w + j - 1);
    return __fsub_rn(hi, lo);
  }
  __device__ void next() { ++j; if (j >= jend) advance_group(); }
};

__device__ __forceinline__ int64_t list_len(const DevGraph& g, int64_t row, const ETypes2& et) {
  if (row < 0) return 0;
  int64_t n = 0;
  for (int32_t k = 0; k < et.K; ++k) {
    int32_t t = et.v[k];
    if (t >= 0 && t < g.T) n += g.grp_ptr[row * g.T + t + 1] - g.grp_ptr[row * g.T + t];
  }
  return n;
}

struct WalkState {
  long long* cur;       // [B] current node id
  long long* parent;    // [B] parent id
  int64_t* cur_row;     // [B] row of cur (-1 absent)
  int64_t* parent_row;  // [B] row whose list is the parent list (-1 = empty list)
};

__global__ void k_walk_init(const long long* __restrict__ nodes, int64_t B, int32_t L, WalkState s,
                            long long* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  long long id = nodes[i];
  out[i * (L + 1)] = id;
  s.cur[i] = id;
  s.parent[i] = id;       // step 0: parent = the start node itself, empty parent list (:272-287)
  s.parent_row[i] = -1;
}

// liveness: the walker draws this step iff its current node has a non-empty neighbor list
__global__ void k_walk_live(DevGraph g, int64_t B, ETypes2 et, WalkState s, uint8_t* live) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  int64_t row = lookup_row(g, (unsigned long long)s.cur[i]);
  s.cur_row[i] = row;
  live[i] = list_len(g, row, et) > 0 ? 1 : 0;
}

// One thread walks the merge (BuildWeights :140-168) and the running f32 prefix
// (CompactWeightedCollection::Init :82-97).  Returns the total; if pick >= 0.0 also returns in *sel
// the first position whose prefix exceeds `pick` (RandomSelect closed form), else the last position.
__device__ float biased_prefix(const DevGraph& g, int64_t crow, const ETypes2& cet, int64_t prow,
                               const ETypes2& pet, long long parent_id, float p, float q, bool select,
                               double pick, long long* sel_id) {
  NbIter c, pn;
  c.init(&g, crow, &cet);
  pn.init(&g, prow, &pet);
  float sum = 0.f;
  long long last_id = 0;
  bool found = false;
  long long pk = pn.done() ? 0 : pn.id();
  while (!c.done()) {
    const long long cid = c.id();
    float w = c.w();
    // advance the parent pointer past smaller ids ("else ++k")
    while (!pn.done() && cid > pk) { pn.next(); if (!pn.done()) pk = pn.id(); }
    if (!pn.done() && cid == pk) {
      pn.next(); if (!pn.done()) pk = pn.id();          // shared neighbor: d_tx = 1, weight unchanged
    } else {
      w = cid != parent_id ? __fdiv_rn(w, q) : __fdiv_rn(w, p);  // d_tx = 2 / d_tx = 0
    }
    sum = __fadd_rn(sum, w);
    last_id = cid;
    if (select && !found && (double)sum > pick) { *sel_id = cid; found = true; if (select) break; }
    c.next();
  }
  if (select && !found) *sel_id = last_id;
  return sum;
}

// Note on the merge above: the reference loop is
//   while (j < nc && k < np) { if (c[j] < p[k]) {bias; ++j} else if (c[j] == p[k]) {++k; ++j} else ++k }
//   while (j < nc) {bias; ++j}
// For a fixed j the inner "else ++k" steps are exactly the `while (cid > pk)` loop, then one of the
// first two branches fires (or k ran out and the tail loop biases) -- same visit order, same state.

__global__ void __launch_bounds__(128) k_walk_step(DevGraph g, int64_t B, int32_t L, int32_t step, ETypes2 cet,
                                                   ETypes2 pet, float p, float q, long long default_node,
                                                   WalkState s, const uint8_t* __restrict__ live,
                                                   const uint32_t* __restrict__ state, bool philox,
                                                   unsigned long long key, long long* __restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  const long long cur = s.cur[i];
  const int64_t crow = s.cur_row[i];
  long long next = default_node;
  if (live[i]) {
    const int64_t prow = s.parent_row[i];
    const long long parent_id = s.parent[i];
    const float total = biased_prefix(g, crow, cet, prow, pet, parent_id, p, q, false, 0.0, nullptr);
    double u, u2;
    if (philox) {
      philox_uniform2((unsigned long long)i, (uint32_t)step, 0x77616C6Bu, key, u, u2);
    } else {
      uint32_t x = state[i];
      u = minstd_uniform(x);
    }
    const double r = pick_r(u, 0.f, total);
    biased_prefix(g, crow, cet, prow, pet, parent_id, p, q, true, r, &next);
  }
  out[i * (L + 1) + step + 1] = next;
  // parent_neighbors_ = this step's lists, parent_ids_ = this step's nodes (:128-131)
  s.parent[i] = cur;
  s.parent_row[i] = crow;
  s.cur[i] = next;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative exact step for the common case: ONE edge type per step and adjacency groups sorted by
// neighbor id (DevGraph::adj_sorted; the reference's merge assumes sorted lists as well).  For sorted
// multisets C (children) and P (parent's list) the two-pointer merge of BuildWeights (:140-168) marks the
// m-th copy (m = 0,1,..) of value v in C as "shared" iff P holds more than m copies of v; everything else is
// biased (/p if it is the parent id, /q otherwise).  That is a per-element predicate, so 32 lanes classify 32
// children at once by streaming P in 32-wide chunks next to C; only the f32 prefix sum (CompactWeightedCollection
// ::Init, compact_weighted_collection.h:82-97) stays serial -- it is evaluated left to right through shuffles,
// ~5 cycles per neighbor, which is the floor for a bit-exact sum.  Pass 1 = total, pass 2 = select.
struct WarpWalk {
  const DevGraph* g;
  int64_t cb, ce, cbase;   // child group [cb, ce), first edge of the child's row
  int64_t pb, pe;          // parent group [pb, pe) (empty at step 0 / for dead parents)
  long long parent_id;
  float p, q;
  int lane;

  __device__ float pass(bool select, double pick, long long* sel) const {
    const unsigned FULL = 0xffffffffu;
    const long long BIG = 0x7fffffffffffffffLL;
    float run = 0.f;                // running prefix, replicated in every lane
    int64_t pk = pb;                // first parent chunk that can still hold values >= the current child minimum
    long long prev_val = 0;
    int prev_run = 0;               // copies of prev_val seen so far at the end of the previous child chunk
    long long last_id = 0;
    for (int64_t cj = cb; cj < ce; cj += 32) {
      const int nvalid = (int)min((int64_t)32, ce - cj);
      const int64_t j = cj + lane;
      const bool valid = lane < nvalid;
      const long long cv = valid ? (long long)__ldg(g->nbr + j) : BIG;
      float w = 0.f;
      if (valid) {
        const float hi = __ldg(g->cum_w + j);
        const float lo = j == cbase ? 0.f : __ldg(g->cum_w + j - 1);
        w = __fsub_rn(hi, lo);
      }
      // m = copies of cv that precede this one in C
      const unsigned peers = __match_any_sync(FULL, cv);
      int m = __popc(peers & ((1u << lane) - 1u));
      if (prev_run && cv == prev_val) m += prev_run;
      // cnt = copies of cv in P: walk the parent chunks whose value range meets [vmin, vmax]
      const long long vmin = __shfl_sync(FULL, cv, 0), vmax = __shfl_sync(FULL, cv, nvalid - 1);
      int cnt = 0;
      for (int64_t tk = pk; tk < pe; tk += 32) {
        const long long pv = tk + lane < pe ? (long long)__ldg(g->nbr + tk + lane) : BIG;
        const long long pmin = __shfl_sync(FULL, pv, 0);
        const long long pmax = __shfl_sync(FULL, pv, (int)min((int64_t)31, pe - tk - 1));
        if (pmax < vmin) { pk = tk + 32; continue; }   // below every remaining child value: never needed again
        if (pmin > vmax) break;                         // above this child chunk: the next chunk restarts at pk
#pragma unroll 8
        for (int k = 0; k < 32; ++k) cnt += (__shfl_sync(FULL, pv, k) == cv) ? 1 : 0;
        if (pmax > vmax) break;
      }
      const bool shared = m < cnt;
      if (valid && !shared) w = cv != parent_id ? __fdiv_rn(w, q) : __fdiv_rn(w, p);   // d_tx = 2 / d_tx = 0
      // serial f32 prefix over the chunk, left to right
      float mine = 0.f;
      for (int k = 0; k < nvalid; ++k) {
        run = __fadd_rn(run, __shfl_sync(FULL, w, k));
        if (lane == k) mine = run;
      }
      if (select) {
        const unsigned hit = __ballot_sync(FULL, valid && (double)mine > pick);
        if (hit) { *sel = __shfl_sync(FULL, cv, __ffs(hit) - 1); return run; }
      }
      last_id = __shfl_sync(FULL, cv, nvalid - 1);
      prev_val = last_id;
      prev_run = __shfl_sync(FULL, m, nvalid - 1) + 1;
    }
    if (select) *sel = last_id;   // RandomSelect's fall-through: the last entry
    return run;
  }
};

__global__ void __launch_bounds__(256) k_walk_step_warp(DevGraph g, int64_t B, int32_t L, int32_t step, int32_t ctype,
                                                        int32_t ptype, float p, float q, long long default_node,
                                                        WalkState s, const uint8_t* __restrict__ live,
                                                        const uint32_t* __restrict__ state, bool philox,
                                                        unsigned long long key, long long* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (i >= B) return;
  const long long cur = s.cur[i];
  const int64_t crow = s.cur_row[i];
  long long next = default_node;
  if (live[i]) {
    const int64_t prow = s.parent_row[i];
    WarpWalk ww;
    ww.g = &g; ww.lane = lane; ww.p = p; ww.q = q; ww.parent_id = s.parent[i];
    ww.cbase = g.grp_ptr[crow * g.T];
    ww.cb = g.grp_ptr[crow * g.T + ctype];
    ww.ce = g.grp_ptr[crow * g.T + ctype + 1];
    ww.pb = ww.pe = 0;
    if (prow >= 0 && ptype >= 0 && ptype < g.T) { ww.pb = g.grp_ptr[prow * g.T + ptype]; ww.pe = g.grp_ptr[prow * g.T + ptype + 1]; }
    const float total = ww.pass(false, 0.0, nullptr);
    double u, u2;
    if (philox) {
      philox_uniform2((unsigned long long)i, (uint32_t)step, 0x77616C6Bu, key, u, u2);
    } else {
      uint32_t x = state[i];
      u = minstd_uniform(x);
    }
    ww.pass(true, pick_r(u, 0.f, total), &next);
  }
  if (lane == 0) {
    out[i * (L + 1) + step + 1] = next;
    s.parent[i] = cur;
    s.parent_row[i] = crow;
    s.cur[i] = next;
  }
}

__global__ void k_walk_col(const unsigned long long* __restrict__ eng, int64_t B, int32_t L, int32_t col,
                           long long default_node, long long* __restrict__ out) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= B) return;
  unsigned long long v = eng[i];
  out[i * (L + 1) + col] = v == 0ull ? default_node : (long long)v;
}

}  // namespace eu

using namespace eu;

extern "C" int eu_random_walk(eu_ctx* c, const int64_t* nodes, int64_t B, const int32_t* etypes, int32_t K,
                              int32_t L, float p, float q, int64_t default_node, int64_t* out) {
  if (!c || B < 0 || L < 0 || K < 0 || K > EU_MAX_ETYPES || (L > 0 && K > 0 && !etypes) || (B > 0 && (!nodes || !out))) {
    set_error("eu_random_walk: bad argument");
    return EU_ERR_INVALID;
  }
  EU_CUDA(cudaSetDevice(c->g->device));
  if (B == 0) return EU_OK;
  const DevGraph& d = c->g->d;
  cudaStream_t s = c->stream;
  const int tb = 256;
  int rc = ctx_reserve(c, hop_scratch_rows(1, B), hop_table_slots(1, B));
  if (rc) return rc;
  const float kEps = 1.0e-6f;
  if (fabs((double)p - 1.0) <= kEps && fabs((double)q - 1.0) <= kEps) {
    // TraditionalRandomWalk: chained sampleNB(count=1); the frontier is the ENGINE id (0 placeholder)
    EU_CUDA(cudaMemcpy2DAsync(out, sizeof(int64_t) * (L + 1), nodes, sizeof(int64_t), sizeof(int64_t), (size_t)B,
                              cudaMemcpyDeviceToDevice, s));
    const unsigned long long* seeds = (const unsigned long long*)nodes;
    for (int l = 0; l < L; ++l) {
      unsigned long long* eng = c->d_front[l & 1];
      rc = hop(c, seeds, B, etypes + (int64_t)l * K, K, 1, default_node, eng, nullptr, nullptr, nullptr,
               l, l > 0 && c->rng == EU_RNG_MINSTD, l + 1 < L, 1);
      if (rc) return rc;
      k_walk_col<<<(unsigned)ceil_div(B, tb), tb, 0, s>>>(eng, B, L, l + 1, default_node, (long long*)out);
      EU_LAUNCHED();
      seeds = eng;
    }
    return EU_OK;
  }
  // node2vec
  rc = ctx_misc(c, 256 + B * (8 + 8 + 8 + 8));
  if (rc) return rc;
  char* m = (char*)c->d_misc + 256;
  WalkState ws;
  ws.cur = (long long*)m; m += 8 * B;
  ws.parent = (long long*)m; m += 8 * B;
  ws.cur_row = (int64_t*)m; m += 8 * B;
  ws.parent_row = (int64_t*)m;
  k_walk_init<<<(unsigned)ceil_div(B, tb), tb, 0, s>>>((const long long*)nodes, B, L, ws, (long long*)out);
  EU_LAUNCHED();
  ETypes2 pet{};
  pet.K = 0;
  for (int l = 0; l < L; ++l) {
    ETypes2 cet{};
    cet.K = K;
    for (int k = 0; k < K; ++k) cet.v[k] = etypes[(int64_t)l * K + k];
    k_walk_live<<<(unsigned)ceil_div(B, tb), tb, 0, s>>>(d, B, cet, ws, c->d_elig);
    EU_LAUNCHED();
    if (c->rng == EU_RNG_MINSTD) {
      rc = launch_state_scan(c, B, 1);
      if (rc) return rc;
    }
    const bool one_sorted_type = d.adj_sorted && cet.K == 1 && pet.K <= 1 && cet.v[0] >= 0 && cet.v[0] < d.T;
    if (one_sorted_type) {
      k_walk_step_warp<<<(unsigned)ceil_div(B * 32, 256), 256, 0, s>>>(d, B, L, l, cet.v[0], pet.K == 1 ? pet.v[0] : -1, p, q,
                                                                       default_node, ws, c->d_elig, c->d_state,
                                                                       c->rng == EU_RNG_PHILOX, c->seed ^ 0x6E32766563ull,
                                                                       (long long*)out);
    } else {
      k_walk_step<<<(unsigned)ceil_div(B, 128), 128, 0, s>>>(d, B, L, l, cet, pet, p, q, default_node, ws, c->d_elig,
                                                             c->d_state, c->rng == EU_RNG_PHILOX, c->seed ^ 0x6E32766563ull,
                                                             (long long*)out);
    }
    EU_LAUNCHED();
    pet = cet;
  }
  return EU_OK;
}

