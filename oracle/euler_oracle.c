/* TEST INFRASTRUCTURE -- see euler_oracle.h.  Plain-C restatement of the reference's
 * algorithms for the minibatch-construction hot path.  Every function cites the
 * reference file:line (relative to /root/reference) it follows.  Compile with
 * -ffp-contract=off (the reference's x86-64 build has no FMA contraction). */
#include "euler_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ RNG */
/* std::minstd_rand0: x <- 16807 x mod (2^31-1); seed(s): x = s mod m, 1 if 0. */
#define EO_M 2147483647ULL
#define EO_A 16807ULL

void eo_seed(eo_rng* r, uint64_t seed) {
  r->x = seed % EO_M;
  if (r->x == 0) r->x = 1;
  r->draws = 0;
}

static inline uint64_t eo_next(eo_rng* r) {
  r->x = (r->x * EO_A) % EO_M;
  return r->x;
}

/* libstdc++ 13 std::generate_canonical<double,53>(minstd_rand0): k = 2 engine calls,
 * R = max-min+1 = 2147483646; sum = (x1-1) + (x2-1)*R; ret = sum / (R*R); ret >= 1 -> nextafter(1,0)
 * (euler/common/random.cc:22-28 uses uniform_real_distribution<double>(0,1),
 * whose operator() is generate_canonical * (b-a) + a = ret * 1.0 + 0.0) */
double eo_uniform(eo_rng* r) {
  const double R = 2147483646.0;
  double sum = (double)(eo_next(r) - 1);
  double tmp = R;
  sum += (double)(eo_next(r) - 1) * tmp;
  tmp *= R;
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  r->draws++;
  return ret * (1.0 - 0.0) + 0.0;
}

static eo_rng g_rng = {1, 0};
void eo_global_seed(uint64_t seed) { eo_seed(&g_rng, seed); }
uint64_t eo_global_draws(void) { return g_rng.draws; }
double eo_global_uniform(void) { return eo_uniform(&g_rng); }
/* save / restore the global engine (lets one process simulate several shard engines) */
uint64_t eo_global_state(void) { return g_rng.x; }
void eo_global_set_state(uint64_t x, uint64_t draws) { g_rng.x = x; g_rng.draws = draws; }

/* ------------------------------------------------------------------ graph */
struct eo_graph {
  int64_t n;
  int32_t T;
  const uint64_t* ids;
  const int32_t* node_type;
  const float* node_w;
  const int64_t* grp_ptr;
  const uint64_t* nbr;
  const float* cum_w;
  const float* grp_cum;
  int32_t feat_dim;
  const float* feat;
  /* id -> row open-addressing table (stands in for unordered_map<NodeID,Node*>, graph.h:87-93) */
  uint64_t cap;
  uint64_t* hkey;
  int64_t* hval;
  /* ids[r] == id_base + r * id_stride for all r (synthetic graphs, shards of them): the lookup is arithmetic and no
   * table is built (a 100M-node table would take minutes to fill; the answer is the same) */
  int dense;
  uint64_t id_base, id_stride;
};

static inline uint64_t eo_mix(uint64_t k) {
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL; k ^= k >> 33;
  return k;
}

eo_graph* eo_graph_create(int64_t n, int32_t T, const uint64_t* ids, const int32_t* node_type,
                          const float* node_w, const int64_t* grp_ptr, const uint64_t* nbr,
                          const float* cum_w, const float* grp_cum, int32_t feat_dim,
                          const float* feat) {
  eo_graph* g = (eo_graph*)calloc(1, sizeof(eo_graph));
  g->n = n; g->T = T; g->ids = ids; g->node_type = node_type; g->node_w = node_w;
  g->grp_ptr = grp_ptr; g->nbr = nbr; g->cum_w = cum_w; g->grp_cum = grp_cum;
  g->feat_dim = feat_dim; g->feat = feat;
  if (n > 1 && ids[1] > ids[0]) {
    const uint64_t st = ids[1] - ids[0];
    int dense = 1;
    for (int64_t r = 2; r < n && dense; ++r) dense = ids[r] == ids[0] + (uint64_t)r * st;
    if (dense) { g->dense = 1; g->id_base = ids[0]; g->id_stride = st; return g; }
  }
  uint64_t cap = 16;
  while (cap < (uint64_t)n * 2) cap <<= 1;
  g->cap = cap;
  g->hkey = (uint64_t*)malloc(cap * sizeof(uint64_t));
  g->hval = (int64_t*)malloc(cap * sizeof(int64_t));
  for (uint64_t i = 0; i < cap; ++i) g->hval[i] = -1;
  for (int64_t r = 0; r < n; ++r) {
    uint64_t h = eo_mix(ids[r]) & (cap - 1);
    while (g->hval[h] >= 0 && g->hkey[h] != ids[r]) h = (h + 1) & (cap - 1);
    g->hkey[h] = ids[r];
    g->hval[h] = r; /* later insert of a duplicate id overwrites, like node_map_[id] = n (graph.cc:164) */
  }
  return g;
}

void eo_graph_destroy(eo_graph* g) {
  if (!g) return;
  free(g->hkey); free(g->hval); free(g);
}

int64_t eo_graph_row(const eo_graph* g, uint64_t id) {
  if (g->dense) {
    if (id < g->id_base || (id - g->id_base) % g->id_stride) return -1;
    const uint64_t r = (id - g->id_base) / g->id_stride;
    return r < (uint64_t)g->n ? (int64_t)r : -1;
  }
  uint64_t h = eo_mix(id) & (g->cap - 1);
  while (g->hval[h] >= 0) {
    if (g->hkey[h] == id) return g->hval[h];
    h = (h + 1) & (g->cap - 1);
  }
  return -1;
}

/* Node::Init (node.cc:46-70): one running f32 sum over all groups of the node; per-group f32
 * sums feed edge_group_collection.Init whose sum_weights_ is again a running f32 sum
 * (compact_weighted_collection.h:82-97). */
void eo_build_cum(int64_t n, int32_t T, const int64_t* grp_ptr, const float* w, float* cum_w,
                  float* grp_cum) {
  for (int64_t r = 0; r < n; ++r) {
    float sum_weight = 0;
    float cwc_sum = 0;
    for (int32_t t = 0; t < T; ++t) {
      float type_weight = 0;
      for (int64_t j = grp_ptr[r * T + t]; j < grp_ptr[r * T + t + 1]; ++j) {
        sum_weight += w[j];
        type_weight += w[j];
        cum_w[j] = sum_weight;
      }
      cwc_sum += type_weight;
      grp_cum[r * T + t] = cwc_sum;
    }
  }
}

/* ------------------------------------------------------------- primitives */
/* RandomSelect, compact_weighted_collection.h:30-52, literal. */
int64_t eo_random_select(const float* sum_weights, int64_t begin_pos, int64_t end_pos, eo_rng* rng) {
  float limit_begin = begin_pos == 0 ? 0 : sum_weights[begin_pos - 1];
  float limit_end = sum_weights[end_pos];
  double r = eo_uniform(rng) * (limit_end - limit_begin) + limit_begin;
  /* size_t low/high in the reference: `high = mid - 1` with mid == 0 wraps to SIZE_MAX and ends the loop */
  uint64_t low = (uint64_t)begin_pos, high = (uint64_t)end_pos, mid = 0;
  int finish = 0;
  while (low <= high && !finish) {
    mid = (low + high) / 2;
    float interval_begin = mid == 0 ? 0 : sum_weights[mid - 1];
    float interval_end = sum_weights[mid];
    if (interval_begin <= r && r < interval_end) {
      finish = 1;
    } else if (interval_begin > r) {
      if (mid == 0) break; /* size_t wrap: low <= SIZE_MAX stays true in the reference only if low..; see note */
      high = mid - 1;
    } else if (interval_end <= r) {
      low = mid + 1;
    }
  }
  return (int64_t)mid;
}
/* note on `mid == 0`: interval_begin is then the literal 0 and r >= 0 always, so the
 * `interval_begin > r` branch cannot be taken with mid == 0; the break is unreachable and only
 * guards the unsigned wrap. */

/* Closed form used by the CUDA kernels: for non-decreasing cum the binary search above returns
 * min(end, first j in [begin,end] with (double)cum[j] > r).  Checked against the literal form in
 * tests/test_oracle_golden.py incl. zero-width intervals and r >= limit_end. */
int64_t eo_random_select_closed(const float* cum, int64_t begin, int64_t end, eo_rng* rng) {
  float limit_begin = begin == 0 ? 0 : cum[begin - 1];
  float limit_end = cum[end];
  double r = eo_uniform(rng) * (limit_end - limit_begin) + limit_begin;
  int64_t j = begin;
  while (j < end && !((double)cum[j] > r)) ++j;
  return j;
}

/* CompactWeightedCollection<T>::Init (h:82-97) + Sample (h:115-124) */
void eo_cwc_sample(const int64_t* ids, const float* w, int64_t n, int64_t ndraws, eo_rng* rng,
                   int64_t* out_ids, float* out_w) {
  float* sum_weights = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
  float sum_weight = 0.0;
  for (int64_t i = 0; i < n; ++i) { sum_weight += w[i]; sum_weights[i] = sum_weight; }
  for (int64_t d = 0; d < ndraws; ++d) {
    int64_t mid = eo_random_select(sum_weights, 0, n - 1, rng);
    float pre = mid > 0 ? sum_weights[mid - 1] : 0;
    out_ids[d] = ids[mid];
    out_w[d] = sum_weights[mid] - pre;
  }
  free(sum_weights);
}

/* AliasMethod::Init, alias_method.cc:23-63 */
void eo_alias_build(const float* weights, int64_t n, float* prob, int64_t* alias) {
  float* w = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
  int64_t* small = (int64_t*)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
  int64_t* large = (int64_t*)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
  int64_t ns = 0, nl = 0;
  memcpy(w, weights, sizeof(float) * n);
  for (int64_t i = 0; i < n; ++i) { prob[i] = 0; alias[i] = 0; }
  double avg = 1 / (double)n;
  for (int64_t i = 0; i < n; ++i) {
    if (w[i] > avg) large[nl++] = i; else small[ns++] = i;
  }
  while (nl > 0 && ns > 0) {
    int64_t less = small[--ns];
    int64_t more = large[--nl];
    prob[less] = w[less] * (float)(uint64_t)n; /* float * size_t -> float multiply */
    alias[less] = more;
    float t = w[more] + w[less];               /* float + float */
    w[more] = (float)((double)t - avg);        /* float - double -> double, stored to float */
    if (w[more] > avg) large[nl++] = more; else small[ns++] = more;
  }
  while (ns > 0) prob[small[--ns]] = 1.0;
  while (nl > 0) prob[large[--nl]] = 1.0;
  free(w); free(small); free(large);
}

/* FastWeightedCollection<T>::Init, fast_weighted_collection.h:54-74 */
void eo_fwc_build(const float* w, int64_t n, float* prob, int64_t* alias, float* sum_weight) {
  float s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += w[i];
  float* norm = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) norm[i] = w[i] / s;
  eo_alias_build(norm, n, prob, alias);
  free(norm);
  *sum_weight = s;
}

/* AliasMethod::Next / NextLong, alias_method.cc:66-78 */
int64_t eo_alias_next(const float* prob, const int64_t* alias, int64_t n, eo_rng* rng) {
  int64_t column = (int64_t)floor((double)n * eo_uniform(rng));
  int coin = eo_uniform(rng) < prob[column];
  return coin ? column : alias[column];
}

/* ------------------------------------------------------------ api.cc level */
/* Node::__SampleNeighbor, node.cc:98-161.  Indices are node-relative in the reference
 * (vectors per node); here `base` = first edge of the row, so the reference's `mid == 0`
 * is `mid == base` and cum is addressed relative to base. */
int32_t eo_node_sample_neighbor(const eo_graph* g, int64_t row, const int32_t* edge_types,
                                int32_t K, int32_t count, eo_rng* rng, uint64_t* out_ids,
                                float* out_w, int32_t* out_t) {
  const int32_t T = g->T; /* ni.edge_group_collection.GetSize() */
  const int64_t base = g->grp_ptr[row * T];
  const float* cum = g->cum_w + base;     /* ni.neighbors_weight */
  const uint64_t* nb = g->nbr + base;     /* ni.neighbors */
  const float* gcum = g->grp_cum + row * T; /* edge_group_collection.sum_weights_ */
#define GROUPS_IDX(t) ((int32_t)(g->grp_ptr[row * T + (t) + 1] - base)) /* ni.neighbor_groups_idx[t] */
  float sub_cum[64];
  int32_t sub_ids[64];
  float sub_sum = 0;
  int use_sub = (K > 1 && K < T);
  if (use_sub) {
    /* rebuild weighted collection, node.cc:105-121 */
    sub_sum = 0.0;
    for (int32_t i = 0; i < K; ++i) {
      int32_t et = edge_types[i];
      if (et >= 0 && et < T) {
        float pre = et > 0 ? gcum[et - 1] : 0; /* CWC::Get, h:134-147 */
        float wt = gcum[et] - pre;
        sub_ids[i] = et;
        sub_sum += wt;
        sub_cum[i] = sub_sum;
      } else {
        return 0; /* err_vec */
      }
    }
  }
  for (int32_t i = 0; i < count; ++i) {
    int32_t edge_type = 0;
    if (K == 1) {
      edge_type = edge_types[0];
      if (edge_type < 0 || edge_type >= T) return 0;
      int32_t pre_idx = edge_type == 0 ? 0 : GROUPS_IDX(edge_type - 1);
      int32_t cur_idx = GROUPS_IDX(edge_type) - 1;
      if (cur_idx < pre_idx) return 0;
    } else if (use_sub) {
      if (sub_sum == 0) return 0;
      edge_type = sub_ids[eo_random_select(sub_cum, 0, K - 1, rng)];
    } else {
      float sumw = T > 0 ? gcum[T - 1] : 0; /* GetSumWeight(): sum_weight_ == last prefix */
      if (sumw == 0) return 0;
      edge_type = (int32_t)eo_random_select(gcum, 0, T - 1, rng); /* ids_[mid] == mid */
    }
    int32_t interval_idx_begin = edge_type == 0 ? 0 : GROUPS_IDX(edge_type - 1);
    int32_t interval_idx_end = GROUPS_IDX(edge_type) - 1;
    if (interval_idx_end < interval_idx_begin) {
      /* the reference indexes out of bounds here (UB); only reachable via
       * RandomSelect's fall-through onto a zero-weight group.  Defined here as "row is empty". */
      return 0;
    }
    int64_t mid = eo_random_select(cum, interval_idx_begin, interval_idx_end, rng);
    float pre_sum_weight = mid <= 0 ? 0 : cum[mid - 1];
    out_ids[i] = nb[mid];
    out_w[i] = cum[mid] - pre_sum_weight;
    out_t[i] = edge_type;
  }
#undef GROUPS_IDX
  return count;
}

/* euler::SampleNeighbor, api.cc:223-236 */
void eo_sample_neighbor(const eo_graph* g, const uint64_t* ids, int64_t n, const int32_t* etypes,
                        int32_t K, int32_t count, eo_rng* rng, uint64_t* out_ids, float* out_w,
                        int32_t* out_t, int32_t* out_len) {
  for (int64_t i = 0; i < n; ++i) {
    int64_t row = eo_graph_row(g, ids[i]);
    out_len[i] = 0;
    if (row >= 0)
      out_len[i] = eo_node_sample_neighbor(g, row, etypes, K, count, rng, out_ids + i * count,
                                           out_w + i * count, out_t + i * count);
  }
}

/* euler::GetFullNeighbor api.cc:208-221 over Node::__GetFullNeighbor node.cc:176-198 */
int64_t eo_get_full_neighbor(const eo_graph* g, const uint64_t* ids, int64_t n,
                             const int32_t* etypes, int32_t K, int64_t cap, int64_t* out_len,
                             uint64_t* out_ids, float* out_w, int32_t* out_t) {
  int64_t tot = 0;
  const int32_t T = g->T;
  for (int64_t i = 0; i < n; ++i) {
    int64_t row = eo_graph_row(g, ids[i]);
    out_len[i] = 0;
    if (row < 0) continue;
    const int64_t base = g->grp_ptr[row * T];
    for (int32_t k = 0; k < K; ++k) {
      int32_t et = etypes[k];
      if (et >= 0 && et < T) {
        for (int64_t j = g->grp_ptr[row * T + et]; j < g->grp_ptr[row * T + et + 1]; ++j) {
          float pre = j == base ? 0 : g->cum_w[j - 1];
          if (tot < cap) {
            out_ids[tot] = g->nbr[j];
            out_w[tot] = g->cum_w[j] - pre;
            out_t[tot] = et;
          }
          ++tot;
          ++out_len[i];
        }
      }
    }
  }
  return tot;
}

/* ------------------------------------------------------ global node sampler */
struct eo_node_sampler {
  int32_t n_types;
  int64_t* size;      /* per type */
  uint64_t** ids;     /* sampler order */
  float** w;          /* FWC::weights_ (= normalised by node_weight_sums_) */
  float** prob;
  int64_t** alias;
  float* fwc_sum;     /* FWC::sum_weight_ per type */
  float* type_sums;   /* node_weight_sums_ */
  float* type_prob;   /* node_type_collection_ alias tables */
  int64_t* type_alias;
  float type_fwc_sum;
};

/* Graph::BuildGlobalSampler, graph.cc:333-370 */
eo_node_sampler* eo_node_sampler_create(const eo_graph* g, const int64_t* order, int64_t n_order,
                                        int32_t n_types) {
  eo_node_sampler* s = (eo_node_sampler*)calloc(1, sizeof(eo_node_sampler));
  s->n_types = n_types;
  s->size = (int64_t*)calloc(n_types, sizeof(int64_t));
  s->ids = (uint64_t**)calloc(n_types, sizeof(void*));
  s->w = (float**)calloc(n_types, sizeof(void*));
  s->prob = (float**)calloc(n_types, sizeof(void*));
  s->alias = (int64_t**)calloc(n_types, sizeof(void*));
  s->fwc_sum = (float*)calloc(n_types, sizeof(float));
  s->type_sums = (float*)calloc(n_types, sizeof(float));
  s->type_prob = (float*)calloc(n_types, sizeof(float));
  s->type_alias = (int64_t*)calloc(n_types, sizeof(int64_t));
  for (int64_t i = 0; i < n_order; ++i) s->size[g->node_type[order[i]]]++;
  for (int32_t t = 0; t < n_types; ++t) {
    int64_t m = s->size[t] > 0 ? s->size[t] : 1;
    s->ids[t] = (uint64_t*)malloc(sizeof(uint64_t) * m);
    s->w[t] = (float*)malloc(sizeof(float) * m);
    s->prob[t] = (float*)malloc(sizeof(float) * m);
    s->alias[t] = (int64_t*)malloc(sizeof(int64_t) * m);
    s->size[t] = 0;
  }
  for (int64_t i = 0; i < n_order; ++i) {
    int64_t r = order[i];
    int32_t t = g->node_type[r];
    s->ids[t][s->size[t]] = g->ids[r];
    s->w[t][s->size[t]] = g->node_w[r];
    s->size[t]++;
    s->type_sums[t] += g->node_w[r];
  }
  for (int32_t t = 0; t < n_types; ++t) {
    for (int64_t i = 0; i < s->size[t]; ++i) s->w[t][i] /= s->type_sums[t];
    eo_fwc_build(s->w[t], s->size[t], s->prob[t], s->alias[t], &s->fwc_sum[t]);
  }
  eo_fwc_build(s->type_sums, n_types, s->type_prob, s->type_alias, &s->type_fwc_sum);
  return s;
}

void eo_node_sampler_destroy(eo_node_sampler* s) {
  if (!s) return;
  for (int32_t t = 0; t < s->n_types; ++t) { free(s->ids[t]); free(s->w[t]); free(s->prob[t]); free(s->alias[t]); }
  free(s->size); free(s->ids); free(s->w); free(s->prob); free(s->alias); free(s->fwc_sum);
  free(s->type_sums); free(s->type_prob); free(s->type_alias); free(s);
}

int64_t eo_node_sampler_size(const eo_node_sampler* s, int32_t type) { return s->size[type]; }

void eo_node_sampler_export(const eo_node_sampler* s, int32_t type, uint64_t* ids, float* w,
                            float* prob, int64_t* alias) {
  memcpy(ids, s->ids[type], sizeof(uint64_t) * s->size[type]);
  memcpy(w, s->w[type], sizeof(float) * s->size[type]);
  memcpy(prob, s->prob[type], sizeof(float) * s->size[type]);
  memcpy(alias, s->alias[type], sizeof(int64_t) * s->size[type]);
}

/* api.cc:32-37 -> Graph::SampleNode graph.cc:221-245 (int) / 247-275 (vector) */
int64_t eo_sample_node(const eo_node_sampler* s, const int32_t* types, int32_t n_types,
                       int32_t count, eo_rng* rng, uint64_t* out) {
  int64_t n_out = 0;
  if (n_types == 1) {
    int32_t node_type = types[0];
    if (node_type == -1) {
      if (s->type_fwc_sum == 0) return 0;
      for (int32_t i = 0; i < count; ++i) {
        node_type = (int32_t)eo_alias_next(s->type_prob, s->type_alias, s->n_types, rng);
        out[n_out++] = s->ids[node_type][eo_alias_next(s->prob[node_type], s->alias[node_type],
                                                        s->size[node_type], rng)];
      }
    } else {
      if (s->fwc_sum[node_type] == 0) return 0;
      for (int32_t i = 0; i < count; ++i)
        out[n_out++] = s->ids[node_type][eo_alias_next(s->prob[node_type], s->alias[node_type],
                                                        s->size[node_type], rng)];
    }
    return n_out;
  }
  /* vector overload: CWC over the listed types in TYPE-ID order (iterates node_type_collection_) */
  float sub_cum[64];
  int32_t sub_ids[64];
  int32_t m = 0;
  float sum = 0.0;
  for (int32_t t = 0; t < s->n_types; ++t) {
    int in_set = 0;
    for (int32_t k = 0; k < n_types; ++k) in_set |= (types[k] == t);
    if (in_set) { sum += s->type_sums[t]; sub_ids[m] = t; sub_cum[m] = sum; ++m; }
  }
  if (sum > 0) {
    for (int32_t i = 0; i < count; ++i) {
      int32_t node_type = sub_ids[eo_random_select(sub_cum, 0, m - 1, rng)];
      out[n_out++] = s->ids[node_type][eo_alias_next(s->prob[node_type], s->alias[node_type],
                                                      s->size[node_type], rng)];
    }
  }
  return n_out;
}

/* --------------------------------------------------------- tf_euler op level */
/* small id->index map for ID_UNIQUE (id_unique_op.cc:41-66) */
typedef struct { uint64_t cap; uint64_t* key; int32_t* val; } eo_umap;
static void umap_init(eo_umap* m, int64_t n) {
  uint64_t cap = 16;
  while (cap < (uint64_t)n * 2) cap <<= 1;
  m->cap = cap;
  m->key = (uint64_t*)malloc(cap * sizeof(uint64_t));
  m->val = (int32_t*)malloc(cap * sizeof(int32_t));
  for (uint64_t i = 0; i < cap; ++i) m->val[i] = -1;
}
static void umap_free(eo_umap* m) { free(m->key); free(m->val); }
static int32_t* umap_slot(eo_umap* m, uint64_t k) {
  uint64_t h = eo_mix(k) & (m->cap - 1);
  while (m->val[h] >= 0 && m->key[h] != k) h = (h + 1) & (m->cap - 1);
  m->key[h] = k;
  return &m->val[h];
}

/* One sampleNB hop as every query runs it (euler/parser/compiler.cc:76-90):
 * ID_UNIQUE -> API_SAMPLE_NB (+ default fill sample_neighbor_op.cc:135-143) ->
 * IDX_GATHER / DATA_GATHER (idx_gather_op.cc:45-55, data_gather_op.cc:34-46). */
static void engine_sample_nb(const eo_graph* g, const uint64_t* ids, int64_t n,
                             const int32_t* etypes, int32_t K, int32_t count, eo_rng* rng,
                             uint64_t* eng_ids, float* eng_w, int32_t* eng_t) {
  eo_umap m;
  umap_init(&m, n);
  uint64_t* uniq = (uint64_t*)malloc(sizeof(uint64_t) * (n > 0 ? n : 1));
  int32_t* gidx = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
  int32_t cnt = 0;
  for (int64_t i = 0; i < n; ++i) {
    int32_t* s = umap_slot(&m, ids[i]);
    if (*s < 0) { *s = cnt; uniq[cnt++] = ids[i]; }
    gidx[i] = *s;
  }
  int64_t c = count > 0 ? count : 1;
  uint64_t* u_ids = (uint64_t*)malloc(sizeof(uint64_t) * (cnt > 0 ? cnt : 1) * c);
  float* u_w = (float*)malloc(sizeof(float) * (cnt > 0 ? cnt : 1) * c);
  int32_t* u_t = (int32_t*)malloc(sizeof(int32_t) * (cnt > 0 ? cnt : 1) * c);
  int32_t* u_len = (int32_t*)malloc(sizeof(int32_t) * (cnt > 0 ? cnt : 1));
  eo_sample_neighbor(g, uniq, cnt, etypes, K, count, rng, u_ids, u_w, u_t, u_len);
  for (int32_t u = 0; u < cnt; ++u) {
    if (u_len[u] == 0)
      for (int32_t j = 0; j < count; ++j) { u_ids[u * c + j] = 0; u_w[u * c + j] = 0; u_t[u * c + j] = 0; }
  }
  for (int64_t i = 0; i < n; ++i) {
    memcpy(eng_ids + i * count, u_ids + gidx[i] * c, sizeof(uint64_t) * count);
    memcpy(eng_w + i * count, u_w + gidx[i] * c, sizeof(float) * count);
    memcpy(eng_t + i * count, u_t + gidx[i] * c, sizeof(int32_t) * count);
  }
  free(uniq); free(gidx); free(u_ids); free(u_w); free(u_t); free(u_len);
  umap_free(&m);
}

/* TF dense packing, tf_euler/kernels/sample_neighbor_op.cc:79-81,114-122 */
static void tf_pack(const uint64_t* e_ids, const float* e_w, const int32_t* e_t, int64_t rows,
                    int32_t count, int64_t default_node, int64_t* out_ids, float* out_w,
                    int32_t* out_t) {
  for (int64_t i = 0; i < rows; ++i) {
    int keep = count > 0 && e_ids[i * count] != 0; /* DEFAULT_UINT64, data_types.cc:27 */
    for (int32_t j = 0; j < count; ++j) {
      int64_t o = i * count + j;
      out_ids[o] = keep ? (int64_t)e_ids[o] : default_node;
      out_w[o] = keep ? e_w[o] : 0.0f;
      out_t[o] = keep ? e_t[o] : -1;
    }
  }
}

static void op_sample_fanout(const eo_graph* g, const int64_t* nodes, int64_t n,
                             const int32_t* etypes, int32_t K, const int32_t* counts, int32_t L,
                             int64_t default_node, eo_rng* rng, int64_t** out_ids, float** out_w,
                             int32_t** out_t) {
  int64_t rows = n;
  uint64_t* seeds = (uint64_t*)malloc(sizeof(uint64_t) * (n > 0 ? n : 1));
  memcpy(seeds, nodes, sizeof(uint64_t) * n);
  for (int32_t l = 0; l < L; ++l) {
    int32_t c = counts[l];
    int64_t m = rows * c > 0 ? rows * c : 1;
    uint64_t* e_ids = (uint64_t*)malloc(sizeof(uint64_t) * m);
    float* e_w = (float*)malloc(sizeof(float) * m);
    int32_t* e_t = (int32_t*)malloc(sizeof(int32_t) * m);
    engine_sample_nb(g, seeds, rows, etypes + l * K, K, c, rng, e_ids, e_w, e_t);
    if (out_ids) tf_pack(e_ids, e_w, e_t, rows, c, default_node, out_ids[l], out_w[l], out_t[l]);
    free(seeds); free(e_w); free(e_t);
    seeds = e_ids; /* sample_fanout_op.cc:36-43: next hop consumes the ENGINE ids (0 placeholders) */
    rows *= c;
  }
  free(seeds);
}

void eo_op_sample_neighbor(const eo_graph* g, const int64_t* nodes, int64_t n,
                           const int32_t* etypes, int32_t K, int32_t count, int64_t default_node,
                           int64_t* out_ids, float* out_w, int32_t* out_t) {
  op_sample_fanout(g, nodes, n, etypes, K, &count, 1, default_node, &g_rng, &out_ids, &out_w, &out_t);
}

void eo_op_sample_fanout(const eo_graph* g, const int64_t* nodes, int64_t n, const int32_t* etypes,
                         int32_t K, const int32_t* counts, int32_t L, int64_t default_node,
                         int64_t** out_ids, float** out_w, int32_t** out_t) {
  op_sample_fanout(g, nodes, n, etypes, K, counts, L, default_node, &g_rng, out_ids, out_w, out_t);
}

/* BuildWeights, tf_euler/kernels/random_walk_op.cc:140-168 (int64 compares) */
static void build_weights(const int64_t* pn, int64_t npn, const int64_t* cn, int64_t ncn,
                          int64_t parent_id, float p, float q, float* w) {
  int64_t j = 0, k = 0;
  while (j < ncn && k < npn) {
    if (cn[j] < pn[k]) {
      if (cn[j] != parent_id) w[j] /= q; else w[j] /= p;
      ++j;
    } else if (cn[j] == pn[k]) {
      ++k; ++j;
    } else {
      ++k;
    }
  }
  while (j < ncn) {
    if (cn[j] != parent_id) w[j] /= q; else w[j] /= p;
    ++j;
  }
}

/* tf_euler.random_walk, tf_euler/kernels/random_walk_op.cc:83-138 (node2vec) and :207-247 (p=q=1) */
void eo_op_random_walk(const eo_graph* g, const int64_t* nodes, int64_t n, const int32_t* etypes,
                       int32_t K, int32_t L, float p, float q, int64_t default_node, int64_t* out) {
  eo_rng* rng = &g_rng;
  for (int64_t i = 0; i < n; ++i) out[i * (L + 1)] = nodes[i];
  const float kEps = 1.0e-6;
  if (fabs(p - 1.0) <= kEps && fabs(q - 1.0) <= kEps) {
    uint64_t* seeds = (uint64_t*)malloc(sizeof(uint64_t) * (n > 0 ? n : 1));
    uint64_t* e_ids = (uint64_t*)malloc(sizeof(uint64_t) * (n > 0 ? n : 1));
    float* e_w = (float*)malloc(sizeof(float) * (n > 0 ? n : 1));
    int32_t* e_t = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    memcpy(seeds, nodes, sizeof(uint64_t) * n);
    for (int32_t l = 0; l < L; ++l) {
      engine_sample_nb(g, seeds, n, etypes + l * K, K, 1, rng, e_ids, e_w, e_t);
      for (int64_t i = 0; i < n; ++i)
        out[i * (L + 1) + l + 1] = e_ids[i] == 0 ? default_node : (int64_t)e_ids[i];
      memcpy(seeds, e_ids, sizeof(uint64_t) * n);
    }
    free(seeds); free(e_ids); free(e_w); free(e_t);
    return;
  }
  /* per-walker parent state */
  int64_t** pn = (int64_t**)calloc(n > 0 ? n : 1, sizeof(int64_t*));
  int64_t* npn = (int64_t*)calloc(n > 0 ? n : 1, sizeof(int64_t));
  int64_t* parent_ids = (int64_t*)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
  int64_t* cur = (int64_t*)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
  memcpy(parent_ids, nodes, sizeof(int64_t) * n);
  memcpy(cur, nodes, sizeof(int64_t) * n);
  for (int32_t step = 0; step < L; ++step) {
    const int32_t* et = etypes + step * K;
    for (int64_t i = 0; i < n; ++i) {
      uint64_t id = (uint64_t)cur[i];
      int64_t len = 0;
      int64_t tot = eo_get_full_neighbor(g, &id, 1, et, K, 0, &len, NULL, NULL, NULL);
      int64_t m = tot > 0 ? tot : 1;
      uint64_t* c_ids = (uint64_t*)malloc(sizeof(uint64_t) * m);
      float* w = (float*)malloc(sizeof(float) * m);
      int32_t* c_t = (int32_t*)malloc(sizeof(int32_t) * m);
      eo_get_full_neighbor(g, &id, 1, et, K, tot, &len, c_ids, w, c_t);
      int64_t sample_id = default_node;
      if (tot > 0) {
        build_weights(pn[i], npn[i], (const int64_t*)c_ids, tot, parent_ids[i], p, q, w);
        int64_t sid; float sw;
        eo_cwc_sample((const int64_t*)c_ids, w, tot, 1, rng, &sid, &sw);
        sample_id = sid;
      }
      out[i * (L + 1) + step + 1] = sample_id;
      free(pn[i]);
      pn[i] = (int64_t*)c_ids; npn[i] = tot; /* parent_neighbors_ = neighbors (:128) */
      parent_ids[i] = cur[i];                /* parent_ids_ <- this query's "nodes" (:129-131) */
      cur[i] = sample_id;
      free(w); free(c_t);
    }
  }
  for (int64_t i = 0; i < n; ++i) free(pn[i]);
  free(pn); free(npn); free(parent_ids); free(cur);
}

/* GetDenseFeature, tf_euler/kernels/get_dense_feature_op.cc:63-121 over api.cc:63-78, one slot of
 * length feat_dim per node; zero fill, missing node -> zeros; copy clipped to dim (see ref_shim). */
void eo_op_get_dense_feature(const eo_graph* g, const int64_t* nodes, int64_t n, int32_t dim,
                             float* out) {
  memset(out, 0, sizeof(float) * n * dim);
  for (int64_t i = 0; i < n; ++i) {
    int64_t row = eo_graph_row(g, (uint64_t)nodes[i]);
    if (row < 0) continue;
    int32_t len = g->feat_dim < dim ? g->feat_dim : dim;
    memcpy(out + i * dim, g->feat + row * g->feat_dim, sizeof(float) * len);
  }
}

/* ------------------------------------------------------------------ mp ops */
/* GatherOp, tf_euler/kernels/gather_op.cc:42-51 (64-bit offsets here; SURVEY A-16) */
void eo_gather(const float* params, int64_t D, const int32_t* idx, int64_t E, float* out) {
  for (int64_t i = 0; i < E; ++i) memcpy(out + i * D, params + (int64_t)idx[i] * D, D * sizeof(float));
}
/* ScatterAddOp, scatter_op.cc:44-55: zero init, serial adds in i order */
void eo_scatter_add(const float* upd, int64_t D, const int32_t* idx, int64_t E, int64_t size,
                    float* out) {
  for (int64_t i = 0; i < size * D; ++i) out[i] = 0;
  for (int64_t i = 0; i < E; ++i)
    for (int64_t j = 0; j < D; ++j) out[(int64_t)idx[i] * D + j] += upd[i * D + j];
}
/* ScatterMaxOp, scatter_op.cc:77-91: init -1e9, strict > */
void eo_scatter_max(const float* upd, int64_t D, const int32_t* idx, int64_t E, int64_t size,
                    float* out) {
  for (int64_t i = 0; i < size * D; ++i) out[i] = -1e9;
  for (int64_t i = 0; i < E; ++i)
    for (int64_t j = 0; j < D; ++j) {
      int64_t o = (int64_t)idx[i] * D + j;
      if (upd[i * D + j] > out[o]) out[o] = upd[i * D + j];
    }
}
/* scatter_mean, tf_euler/python/euler_ops/mp_ops.py:65-69: add / (add(ones) + 1e-7) in f32 */
void eo_scatter_mean(const float* upd, int64_t D, const int32_t* idx, int64_t E, int64_t size,
                     float* out) {
  float* cnt = (float*)calloc(size > 0 ? size : 1, sizeof(float));
  eo_scatter_add(upd, D, idx, E, size, out);
  for (int64_t i = 0; i < E; ++i) cnt[idx[i]] += 1.0f;
  for (int64_t r = 0; r < size; ++r) {
    float c = cnt[r] + 1e-7f;
    for (int64_t j = 0; j < D; ++j) out[r * D + j] = out[r * D + j] / c;
  }
  free(cnt);
}

/* euler/core/kernels/id_split_op.cc:46-49 */
int32_t eo_shard_of(uint64_t id, int32_t num_partitions, int32_t shard_num) {
  return (int32_t)((id % (uint64_t)num_partitions) % (uint64_t)shard_num);
}

/* -------------------------------------------------------------- CPU baseline */
typedef struct {
  const eo_graph* g; const int64_t* seeds; int64_t n_batches, B; const int32_t* etypes; int32_t K;
  const int32_t* counts; int32_t L; int32_t iters; int32_t tid; int64_t edges;
} eo_bench_arg;

static void* bench_worker(void* p) {
  eo_bench_arg* a = (eo_bench_arg*)p;
  eo_rng rng;
  eo_seed(&rng, 12345 + a->tid);
  int64_t per_batch = 0, rows = a->B;
  for (int32_t l = 0; l < a->L; ++l) { rows *= a->counts[l]; per_batch += rows; }
  int64_t** o_ids = (int64_t**)malloc(sizeof(void*) * a->L);
  float** o_w = (float**)malloc(sizeof(void*) * a->L);
  int32_t** o_t = (int32_t**)malloc(sizeof(void*) * a->L);
  rows = a->B;
  for (int32_t l = 0; l < a->L; ++l) {
    rows *= a->counts[l];
    o_ids[l] = (int64_t*)malloc(sizeof(int64_t) * rows);
    o_w[l] = (float*)malloc(sizeof(float) * rows);
    o_t[l] = (int32_t*)malloc(sizeof(int32_t) * rows);
  }
  for (int32_t b = 0; b < a->iters; ++b) {
    const int64_t* s = a->seeds + (((int64_t)a->tid * a->iters + b) % a->n_batches) * a->B;
    op_sample_fanout(a->g, s, a->B, a->etypes, a->K, a->counts, a->L, -1, &rng, o_ids, o_w, o_t);
    a->edges += per_batch;
  }
  for (int32_t l = 0; l < a->L; ++l) { free(o_ids[l]); free(o_w[l]); free(o_t[l]); }
  free(o_ids); free(o_w); free(o_t);
  return NULL;
}

/* full step (sample_fanout + dense features of every hop + neighbor means), see ref_shim.cc ref_bench_step */
typedef struct {
  const eo_graph* g; const int64_t* seeds; int64_t n_batches, B; const int32_t* etypes; int32_t K;
  const int32_t* counts; int32_t L; int32_t dim; int32_t iters; int32_t tid; int64_t edges;
} eo_step_arg;

static void* step_worker(void* p) {
  eo_step_arg* a = (eo_step_arg*)p;
  eo_rng rng;
  eo_seed(&rng, 12345 + a->tid);
  const int32_t L = a->L, dim = a->dim;
  int64_t rows[17];
  int64_t per_batch = 0;
  rows[0] = a->B;
  for (int32_t l = 0; l < L; ++l) { rows[l + 1] = rows[l] * a->counts[l]; per_batch += rows[l + 1]; }
  int64_t* o_ids[16]; float* o_w[16]; int32_t* o_t[16]; float* feat[17]; float* agg[16];
  for (int32_t l = 0; l < L; ++l) {
    o_ids[l] = (int64_t*)malloc(sizeof(int64_t) * rows[l + 1]);
    o_w[l] = (float*)malloc(sizeof(float) * rows[l + 1]);
    o_t[l] = (int32_t*)malloc(sizeof(int32_t) * rows[l + 1]);
    agg[l] = (float*)malloc(sizeof(float) * rows[l] * dim);
  }
  for (int32_t l = 0; l <= L; ++l) feat[l] = (float*)malloc(sizeof(float) * rows[l] * dim);
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * rows[L]);
  for (int32_t b = 0; b < a->iters; ++b) {
    const int64_t* s = a->seeds + (((int64_t)a->tid * a->iters + b) % a->n_batches) * a->B;
    op_sample_fanout(a->g, s, a->B, a->etypes, a->K, a->counts, L, -1, &rng, o_ids, o_w, o_t);
    for (int32_t l = 0; l <= L; ++l) eo_op_get_dense_feature(a->g, l == 0 ? s : o_ids[l - 1], rows[l], dim, feat[l]);
    for (int32_t l = 0; l < L; ++l) {
      for (int64_t i = 0; i < rows[l + 1]; ++i) idx[i] = (int32_t)(i / a->counts[l]);
      eo_scatter_mean(feat[l + 1], dim, idx, rows[l + 1], rows[l], agg[l]);
    }
    a->edges += per_batch;
  }
  for (int32_t l = 0; l < L; ++l) { free(o_ids[l]); free(o_w[l]); free(o_t[l]); free(agg[l]); }
  for (int32_t l = 0; l <= L; ++l) free(feat[l]);
  free(idx);
  return NULL;
}

double eo_bench_step(const eo_graph* g, const int64_t* seeds, int64_t n_batches, int64_t B,
                     const int32_t* etypes, int32_t K, const int32_t* counts, int32_t L, int32_t dim,
                     int32_t n_threads, int32_t iters, int64_t* edges) {
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  eo_step_arg* args = (eo_step_arg*)calloc(n_threads, sizeof(eo_step_arg));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int32_t t = 0; t < n_threads; ++t) {
    eo_step_arg a = {g, seeds, n_batches, B, etypes, K, counts, L, dim, iters, t, 0};
    args[t] = a;
    pthread_create(&th[t], NULL, step_worker, &args[t]);
  }
  int64_t tot = 0;
  for (int32_t t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); tot += args[t].edges; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  *edges = tot;
  free(th); free(args);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double eo_bench_fanout(const eo_graph* g, const int64_t* seeds, int64_t n_batches, int64_t B,
                       const int32_t* etypes, int32_t K, const int32_t* counts, int32_t L,
                       int32_t n_threads, int32_t iters, int64_t* edges) {
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
  eo_bench_arg* args = (eo_bench_arg*)calloc(n_threads, sizeof(eo_bench_arg));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int32_t t = 0; t < n_threads; ++t) {
    eo_bench_arg a = {g, seeds, n_batches, B, etypes, K, counts, L, iters, t, 0};
    args[t] = a;
    pthread_create(&th[t], NULL, bench_worker, &args[t]);
  }
  int64_t tot = 0;
  for (int32_t t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); tot += args[t].edges; }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  *edges = tot;
  free(th); free(args);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
