/* TEST INFRASTRUCTURE -- CPU restatement ("oracle") of alibaba/euler's
 * minibatch-construction hot path.  NOT part of the shipped product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (euler_b200/) never does.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_vs_ref.py,
 * run where /root/reference exists) against oracle/_ref/libeuler_ref.so, i.e. the
 * reference's own sources compiled unmodified with only random.cc replaced by a
 * seedable equivalent, and against the reference tests' golden vectors
 * (tests/golden/, tests/test_oracle_golden.py).
 *
 * All functions cite the reference file:line they restate (paths relative to
 * /root/reference).
 */
#ifndef EULER_ORACLE_H_
#define EULER_ORACLE_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- RNG: std::default_random_engine (= minstd_rand0) + uniform_real_distribution<double>(0,1)
 * euler/common/random.cc:22-28, libstdc++ 13 generate_canonical<double,53> (2 engine calls / uniform) */
typedef struct { uint64_t x; uint64_t draws; } eo_rng;
void eo_seed(eo_rng* r, uint64_t seed);
double eo_uniform(eo_rng* r);
/* the global stream used by the eo_op_* entry points (one engine, like one reference thread) */
void eo_global_seed(uint64_t seed);
uint64_t eo_global_draws(void);
double eo_global_uniform(void);
uint64_t eo_global_state(void);
void eo_global_set_state(uint64_t x, uint64_t draws);

/* ---- graph in CSR form (what Node/NeighborInfo hold, euler/core/graph/node.h:49-57) */
typedef struct eo_graph eo_graph;
/* Arrays are borrowed (caller keeps them alive).  grp_ptr: [n*T+1] group boundaries;
 * cum_w: per-node global cumulative f32 weights as stored (node.cc:59-65);
 * grp_cum: [n*T] = edge_group_collection.sum_weights_ (compact_weighted_collection.h:82-97) */
eo_graph* eo_graph_create(int64_t n, int32_t T, const uint64_t* ids, const int32_t* node_type,
                          const float* node_w, const int64_t* grp_ptr, const uint64_t* nbr,
                          const float* cum_w, const float* grp_cum, int32_t feat_dim,
                          const float* feat);
void eo_graph_destroy(eo_graph* g);
int64_t eo_graph_row(const eo_graph* g, uint64_t id); /* -1 if absent (Graph::GetNodeByID, graph.h:87-93) */

/* Node::Init prefix accumulation (node.cc:37-70): raw weights -> cum_w, grp_cum */
void eo_build_cum(int64_t n, int32_t T, const int64_t* grp_ptr, const float* w, float* cum_w,
                  float* grp_cum);

/* ---- primitives */
/* RandomSelect (compact_weighted_collection.h:30-52) on cum[0..], indices relative to the array */
int64_t eo_random_select(const float* cum, int64_t begin, int64_t end, eo_rng* r);
/* closed form proved equivalent for non-decreasing cum: min(end, first j in [begin,end] with cum[j] > r) */
int64_t eo_random_select_closed(const float* cum, int64_t begin, int64_t end, eo_rng* r);
/* CompactWeightedCollection<int64>::Init + Sample x ndraws (…h:82-128) */
void eo_cwc_sample(const int64_t* ids, const float* w, int64_t n, int64_t ndraws, eo_rng* r,
                   int64_t* out_ids, float* out_w);
/* AliasMethod::Init (alias_method.cc:23-63) on normalised weights */
void eo_alias_build(const float* norm_w, int64_t n, float* prob, int64_t* alias);
/* FastWeightedCollection::Init normalisation (fast_weighted_collection.h:54-74) + alias build */
void eo_fwc_build(const float* w, int64_t n, float* prob, int64_t* alias, float* sum_weight);
/* AliasMethod::Next (alias_method.cc:66-78) */
int64_t eo_alias_next(const float* prob, const int64_t* alias, int64_t n, eo_rng* r);

/* ---- api.cc level */
/* Node::__SampleNeighbor (node.cc:98-161) for one row; returns 0 or count */
int32_t eo_node_sample_neighbor(const eo_graph* g, int64_t row, const int32_t* etypes, int32_t K,
                                int32_t count, eo_rng* r, uint64_t* out_ids, float* out_w,
                                int32_t* out_t);
/* euler::SampleNeighbor (api.cc:223-236); rows dense at i*count, out_len 0|count */
void eo_sample_neighbor(const eo_graph* g, const uint64_t* ids, int64_t n, const int32_t* etypes,
                        int32_t K, int32_t count, eo_rng* r, uint64_t* out_ids, float* out_w,
                        int32_t* out_t, int32_t* out_len);
/* euler::GetFullNeighbor (api.cc:208-221, node.cc:176-198); two-pass like the shim */
int64_t eo_get_full_neighbor(const eo_graph* g, const uint64_t* ids, int64_t n,
                             const int32_t* etypes, int32_t K, int64_t cap, int64_t* out_len,
                             uint64_t* out_ids, float* out_w, int32_t* out_t);

/* ---- global node sampler (graph.cc:221-275,333-370) */
typedef struct eo_node_sampler eo_node_sampler;
/* order: node rows in the order BuildGlobalSampler visits them (unordered_map iteration order in
 * the reference; an input here) */
eo_node_sampler* eo_node_sampler_create(const eo_graph* g, const int64_t* order, int64_t n_order,
                                        int32_t n_types);
void eo_node_sampler_destroy(eo_node_sampler* s);
int64_t eo_node_sampler_size(const eo_node_sampler* s, int32_t type);
void eo_node_sampler_export(const eo_node_sampler* s, int32_t type, uint64_t* ids, float* w,
                            float* prob, int64_t* alias);
/* euler::SampleNode (api.cc:32-37): 1 type -> Graph::SampleNode(int,int) (-1 = all types),
 * else the vector overload.  Returns number of ids written (0 or count). */
int64_t eo_sample_node(const eo_node_sampler* s, const int32_t* types, int32_t n_types,
                       int32_t count, eo_rng* r, uint64_t* out);

/* ---- tf_euler op level (use the global stream) */
void eo_op_sample_neighbor(const eo_graph* g, const int64_t* nodes, int64_t n,
                           const int32_t* etypes, int32_t K, int32_t count, int64_t default_node,
                           int64_t* out_ids, float* out_w, int32_t* out_t);
void eo_op_sample_fanout(const eo_graph* g, const int64_t* nodes, int64_t n, const int32_t* etypes,
                         int32_t K, const int32_t* counts, int32_t L, int64_t default_node,
                         int64_t** out_ids, float** out_w, int32_t** out_t);
void eo_op_random_walk(const eo_graph* g, const int64_t* nodes, int64_t n, const int32_t* etypes,
                       int32_t K, int32_t L, float p, float q, int64_t default_node, int64_t* out);
/* GetDenseFeature (tf_euler/kernels/get_dense_feature_op.cc:63-121) one feature slot */
void eo_op_get_dense_feature(const eo_graph* g, const int64_t* nodes, int64_t n, int32_t dim,
                             float* out);

/* ---- mp ops (tf_euler/kernels/gather_op.cc:42-51, scatter_op.cc:44-55,77-91; mp_ops.py:65-69) */
void eo_gather(const float* params, int64_t D, const int32_t* idx, int64_t E, float* out);
void eo_scatter_add(const float* upd, int64_t D, const int32_t* idx, int64_t E, int64_t size,
                    float* out);
void eo_scatter_max(const float* upd, int64_t D, const int32_t* idx, int64_t E, int64_t size,
                    float* out);
void eo_scatter_mean(const float* upd, int64_t D, const int32_t* idx, int64_t E, int64_t size,
                     float* out);

/* ---- sharding (euler/core/kernels/id_split_op.cc:46-49) */
int32_t eo_shard_of(uint64_t id, int32_t num_partitions, int32_t shard_num);

/* ---- CPU baseline loops for bench.py (port kind).  n_threads workers, own RNG each. */
double eo_bench_fanout(const eo_graph* g, const int64_t* seeds, int64_t n_batches, int64_t B,
                       const int32_t* etypes, int32_t K, const int32_t* counts, int32_t L,
                       int32_t n_threads, int32_t iters, int64_t* edges);

double eo_bench_step(const eo_graph* g, const int64_t* seeds, int64_t n_batches, int64_t B,
                     const int32_t* etypes, int32_t K, const int32_t* counts, int32_t L, int32_t dim,
                     int32_t n_threads, int32_t iters, int64_t* edges);

#ifdef __cplusplus
}
#endif
#endif /* EULER_ORACLE_H_ */
