// TEST INFRASTRUCTURE (oracle/_ref) -- not part of the shipped product.
//
// A thin extern "C" shim over the UNMODIFIED reference sources compiled where
// they lie under /root/reference (see oracle/Makefile).  It exposes, for the
// minibatch-construction hot path only:
//   * the reference's own graph loader           (euler/core/graph/graph.cc:72-120)
//   * in-memory construction via Node::Init       (euler/core/graph/node.cc:37-96)
//   * euler::SampleNeighbor / GetFullNeighbor /    (euler/core/api/api.cc:208-236)
//     SampleNode / GetNodeFloat32Feature           (euler/core/api/api.cc:32-37,63-78)
//   * read-only export of the loaded graph so fixtures can be generated
// plus restatements of the engine/TF wrappers that cannot be compiled here
// (they need protobuf / TensorFlow), each citing the lines it follows.
//
// Built with -fno-access-control so the shim can read private members of the
// reference classes without editing the reference.
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <malloc.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "euler/common/alias_method.h"
#include "euler/common/compact_weighted_collection.h"
#include "euler/common/fast_weighted_collection.h"
#include "euler/common/server_register.h"
#include "euler/core/api/api.h"
#include "euler/core/graph/graph.h"
#include "euler/core/graph/graph_meta.h"
#include "euler/core/graph/node.h"

extern "C" void ref_seed(uint64_t seed);

namespace euler {
// zookeeper is not built; Graph::DeregisterRemote is never called on this path.
std::shared_ptr<ServerRegister> GetServerRegister(const std::string&,
                                                  const std::string&) {
  return nullptr;
}
}  // namespace euler

namespace {

// The reference is built with jemalloc by default (CMakeLists.txt:13,41-43); it is not available
// offline.  Keep glibc malloc from returning freed pages to the OS on every batch (the feature path
// allocates ~3 small vectors per node), which otherwise serialises worker threads on mmap_sem.
struct MallocTune {
  MallocTune() {
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 256 << 20);
  }
} g_malloc_tune;

euler::Graph& G() { return euler::Graph::Instance(); }

void ResetGraph() {
  euler::Graph* g = &G();
  g->~Graph();
  new (g) euler::Graph();
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- lifecycle
void ref_graph_reset() { ResetGraph(); }

// The reference's own loader, shard 0 of 1 (graph.cc:72-120).
int ref_graph_load(const char* dir, const char* sampler_type,
                   const char* data_type) {
  ResetGraph();
  auto s = G().Init(0, 1, sampler_type, dir, data_type);
  return s.ok() ? 0 : 1;
}

// In-memory build through Node::Init + Graph::AddNode (node.cc:37-96,
// graph.cc:162-166) -- the route BASELINE.md section 3 prescribes for synthetic
// graphs.  w is the RAW edge weight; Node::Init accumulates the f32 prefix.
// grp_ptr has n*T+1 entries; group g of node r spans [grp_ptr[r*T+g], grp_ptr[r*T+g+1]).
int ref_graph_build(int64_t n, const uint64_t* ids, const int32_t* types,
                    const float* node_w, int32_t T, const int64_t* grp_ptr,
                    const uint64_t* nbr, const float* w, int32_t n_node_types,
                    int32_t feat_dim, const float* feat, int32_t build_sampler) {
  ResetGraph();
  euler::Graph& g = G();
  g.reserveNodeMap(n);
  std::vector<std::vector<uint64_t>> nb(T);
  std::vector<std::vector<float>> nw(T);
  std::vector<std::vector<uint64_t>> u64f;
  std::vector<std::vector<float>> f32f(feat_dim > 0 ? 1 : 0);
  std::vector<std::string> binf;
  for (int64_t r = 0; r < n; ++r) {
    for (int32_t t = 0; t < T; ++t) {
      int64_t b = grp_ptr[r * T + t], e = grp_ptr[r * T + t + 1];
      nb[t].assign(nbr + b, nbr + e);
      nw[t].assign(w + b, w + e);
    }
    if (feat_dim > 0) f32f[0].assign(feat + r * feat_dim, feat + (r + 1) * feat_dim);
    euler::Node* node = new euler::Node(ids[r], node_w[r], types[r]);
    if (!node->Init(nb, nw, u64f, f32f, binf)) return 1;
    g.AddNode(node);
  }
  std::unordered_map<std::string, uint32_t> ntm, etm;
  for (int32_t t = 0; t < n_node_types; ++t) ntm[std::to_string(t)] = t;
  for (int32_t t = 0; t < T; ++t) etm[std::to_string(t)] = t;
  euler::FeatureInfoMap nfi, efi;
  if (feat_dim > 0) nfi["dense_feat"] = std::make_tuple(euler::kDense, 0, (int64_t)feat_dim);
  euler::GraphMeta meta("synthetic", "0", n, grp_ptr[n * T], 1, nfi, efi, ntm, etm);
  g.set_meta(meta);
  if (build_sampler) g.BuildGlobalSampler();
  return 0;
}

// ------------------------------------------------------------------- export
int64_t ref_node_count() { return G().getNodeSize(); }
int32_t ref_node_type_num() { return G().GetNodeTypeNum(); }
int32_t ref_edge_type_num() { return G().GetEdgeTypeNum(); }

// Node ids in unordered_map iteration order -- the order BuildGlobalSampler
// walks (graph.cc:349-354).
void ref_export_node_ids(uint64_t* ids) {
  int64_t i = 0;
  for (auto& it : G().node_map_) ids[i++] = it.first;
}

// returns 0 if the node exists
int ref_node_info(uint64_t id, int32_t* type, float* weight, int32_t* n_groups,
                  int32_t* degree, int32_t* n_f32_slots, int32_t* n_f32_vals) {
  euler::Node* n = G().GetNodeByID(id);
  if (n == nullptr) return 1;
  *type = n->GetType();
  *weight = n->GetWeight();
  *n_groups = n->neighbor_info_.neighbor_groups_idx.size();
  *degree = n->neighbor_info_.neighbors.size();
  *n_f32_slots = n->float_features_idx_.size();
  *n_f32_vals = n->float_features_.size();
  return 0;
}

// Raw stored adjacency (NeighborInfo, node.h:49-57): group ends, ids, the
// node-global cumulative weights, and the edge-group CWC's prefix sums.
int ref_node_adj(uint64_t id, int32_t* group_ends, uint64_t* nbr, float* cum_w,
                 float* group_cum) {
  euler::Node* n = G().GetNodeByID(id);
  if (n == nullptr) return 1;
  auto& ni = n->neighbor_info_;
  std::copy(ni.neighbor_groups_idx.begin(), ni.neighbor_groups_idx.end(), group_ends);
  std::copy(ni.neighbors.begin(), ni.neighbors.end(), nbr);
  std::copy(ni.neighbors_weight.begin(), ni.neighbors_weight.end(), cum_w);
  auto& sw = ni.edge_group_collection.sum_weights_;
  std::copy(sw.begin(), sw.end(), group_cum);
  return 0;
}

int ref_node_f32feat(uint64_t id, int32_t* slot_ends, float* vals) {
  euler::Node* n = G().GetNodeByID(id);
  if (n == nullptr) return 1;
  std::copy(n->float_features_idx_.begin(), n->float_features_idx_.end(), slot_ends);
  std::copy(n->float_features_.begin(), n->float_features_.end(), vals);
  return 0;
}

// Global node sampler of one type: ids in sampler order, weights, alias tables
// (fast_weighted_collection.h:54-74, alias_method.cc:23-63).
int64_t ref_sampler_size(int32_t type) {
  if (type < 0 || type >= (int32_t)G().node_samplers_.size()) return -1;
  return G().node_samplers_[type].ids_.size();
}
void ref_sampler_export(int32_t type, uint64_t* ids, float* weights, float* prob,
                        int64_t* alias) {
  auto& s = G().node_samplers_[type];
  std::copy(s.ids_.begin(), s.ids_.end(), ids);
  std::copy(s.weights_.begin(), s.weights_.end(), weights);
  std::copy(s.alias_.prob_.begin(), s.alias_.prob_.end(), prob);
  std::copy(s.alias_.alias_.begin(), s.alias_.alias_.end(), alias);
}
void ref_type_sampler_export(float* type_weight_sums, float* prob, int64_t* alias) {
  auto& s = G().node_type_collection_;
  std::copy(s.weights_.begin(), s.weights_.end(), type_weight_sums);
  std::copy(s.alias_.prob_.begin(), s.alias_.prob_.end(), prob);
  std::copy(s.alias_.alias_.begin(), s.alias_.alias_.end(), alias);
}

// --------------------------------------------------------- sampling primitives
// RandomSelect on a caller-provided cumulative array (compact_weighted_collection.h:30-52)
int64_t ref_random_select(const float* cum, int64_t n, int64_t begin, int64_t end) {
  std::vector<float> v(cum, cum + n);
  return euler::common::RandomSelect<uint64_t>(v, begin, end);
}

// CompactWeightedCollection<int64>::Init + ndraws x Sample (…h:82-128)
void ref_cwc_sample(const int64_t* ids, const float* weights, int64_t n,
                    int64_t ndraws, int64_t* out_ids, float* out_w) {
  euler::common::CompactWeightedCollection<int64_t> c;
  c.Init(std::vector<int64_t>(ids, ids + n), std::vector<float>(weights, weights + n));
  for (int64_t i = 0; i < ndraws; ++i) {
    auto p = c.Sample();
    out_ids[i] = p.first;
    out_w[i] = p.second;
  }
}

// AliasMethod::Init on already-normalised weights + table export (alias_method.cc:23-63)
void ref_alias_build(const float* norm_w, int64_t n, float* prob, int64_t* alias) {
  euler::common::AliasMethod a;
  a.Init(std::vector<float>(norm_w, norm_w + n));
  std::copy(a.prob_.begin(), a.prob_.end(), prob);
  std::copy(a.alias_.begin(), a.alias_.end(), alias);
}

// FastWeightedCollection<uint64>::Init + ndraws x Sample
void ref_fwc_sample(const uint64_t* ids, const float* weights, int64_t n,
                    int64_t ndraws, uint64_t* out_ids) {
  euler::common::FastWeightedCollection<uint64_t> c;
  c.Init(std::vector<uint64_t>(ids, ids + n), std::vector<float>(weights, weights + n));
  for (int64_t i = 0; i < ndraws; ++i) out_ids[i] = c.Sample().first;
}

// --------------------------------------------------------------- api.cc calls
// euler::SampleNeighbor (api.cc:223-236).  out_len[i] is 0 (missing node /
// empty result) or count; rows are written densely at i*count.
void ref_sample_neighbor(const uint64_t* ids, int64_t n, const int32_t* etypes,
                         int32_t K, int32_t count, uint64_t* out_ids, float* out_w,
                         int32_t* out_t, int32_t* out_len) {
  euler::NodeIdVec v(ids, ids + n);
  std::vector<int> et(etypes, etypes + K);
  auto res = euler::SampleNeighbor(v, et, count);
  for (int64_t i = 0; i < n; ++i) {
    out_len[i] = res[i].size();
    for (size_t j = 0; j < res[i].size(); ++j) {
      out_ids[i * count + j] = std::get<0>(res[i][j]);
      out_w[i * count + j] = std::get<1>(res[i][j]);
      out_t[i * count + j] = std::get<2>(res[i][j]);
    }
  }
}

// euler::GetFullNeighbor (api.cc:208-221).  Two-pass: with cap == 0 only
// out_len is filled.  Returns the total number of entries.
int64_t ref_get_full_neighbor(const uint64_t* ids, int64_t n, const int32_t* etypes,
                              int32_t K, int64_t cap, int64_t* out_len,
                              uint64_t* out_ids, float* out_w, int32_t* out_t) {
  euler::NodeIdVec v(ids, ids + n);
  std::vector<int> et(etypes, etypes + K);
  auto res = euler::GetFullNeighbor(v, et);
  int64_t tot = 0;
  for (int64_t i = 0; i < n; ++i) {
    out_len[i] = res[i].size();
    for (auto& iw : res[i]) {
      if (tot < cap) {
        out_ids[tot] = std::get<0>(iw);
        out_w[tot] = std::get<1>(iw);
        out_t[tot] = std::get<2>(iw);
      }
      ++tot;
    }
  }
  return tot;
}

// euler::SampleNode (api.cc:32-37).  Returns number of ids produced.
int64_t ref_sample_node(const int32_t* types, int32_t n_types, int32_t count,
                        uint64_t* out) {
  std::vector<int> t(types, types + n_types);
  auto v = euler::SampleNode(t, count);
  std::copy(v.begin(), v.end(), out);
  return v.size();
}

// euler::SampleEdge (api.cc:39-44 -> Graph::SampleEdge graph.cc:277-331).  out: [count][3] = src, dst, type.
// Returns the number of edges produced (0 for several types / -1: edge_type_collection_ is never initialised upstream).
int64_t ref_sample_edge(const int32_t* types, int32_t n_types, int32_t count, uint64_t* out) {
  std::vector<int> t(types, types + n_types);
  auto v = euler::SampleEdge(t, count);
  for (size_t i = 0; i < v.size(); ++i) {
    out[i * 3] = std::get<0>(v[i]); out[i * 3 + 1] = std::get<1>(v[i]); out[i * 3 + 2] = (uint64_t)(int64_t)std::get<2>(v[i]);
  }
  return v.size();
}

// euler::GetNodeFloat32Feature (api.cc:63-78) for one feature slot, with the TF
// kernel's zero fill (tf_euler/kernels/get_dense_feature_op.cc:66-75,108-115).
// The TF kernel copies each row's TRUE length; here the copy is clipped to dim
// (longer rows are a caller error there: heap overflow) and the true length is
// reported in out_len.
void ref_get_dense_feature(const uint64_t* ids, int64_t n, int32_t fid, int32_t dim,
                           float* out, int32_t* out_len) {
  euler::NodeIdVec v(ids, ids + n);
  std::vector<int> fids(1, fid);
  auto res = euler::GetNodeFloat32Feature(v, fids);
  std::fill(out, out + n * (int64_t)dim, 0.0f);
  for (int64_t i = 0; i < n; ++i) {
    int32_t len = res[i].empty() ? 0 : (int32_t)res[i][0].size();
    out_len[i] = len;
    if (len > 0) std::copy(res[i][0].begin(), res[i][0].begin() + std::min(len, dim), out + i * dim);
  }
}

// --------------------------------------------- engine / TF wrapper restatement
// One sampleNB hop as the query engine runs it in every mode
// (euler/parser/compiler.cc:76-90):
//   ID_UNIQUE (id_unique_op.cc:41-66) -> API_SAMPLE_NB (sample_neighbor_op.cc:37-147,
//   default fill :135-143) -> IDX_GATHER/DATA_GATHER (idx_gather_op.cc:45-55,
//   data_gather_op.cc:34-46).
// eng_* are the engine outputs [n,count] (placeholders (0,0.0,0)).
static void EngineSampleNB(const uint64_t* ids, int64_t n, const std::vector<int>& et,
                           int32_t count, uint64_t* eng_ids, float* eng_w, int32_t* eng_t) {
  std::unordered_map<uint64_t, int32_t> ids_map;
  ids_map.reserve(n);
  std::vector<uint64_t> uniq;
  uniq.reserve(n);
  int32_t cnt = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (ids_map.find(ids[i]) == ids_map.end()) {
      ids_map[ids[i]] = cnt++;
      uniq.push_back(ids[i]);
    }
  }
  auto res = euler::SampleNeighbor(uniq, et, count);
  for (auto& item : res) {
    if (item.empty()) {
      item.reserve(count);
      for (int32_t i = 0; i < count; ++i) item.push_back(euler::IdWeightPair(0, 0, 0));
    }
  }
  for (int64_t i = 0; i < n; ++i) {
    auto& row = res[ids_map.at(ids[i])];
    for (int32_t j = 0; j < count; ++j) {
      eng_ids[i * count + j] = std::get<0>(row[j]);
      eng_w[i * count + j] = std::get<1>(row[j]);
      eng_t[i * count + j] = std::get<2>(row[j]);
    }
  }
}

// TF-level dense packing (tf_euler/kernels/sample_neighbor_op.cc:79-81,114-122):
// pre-fill (default_node, 0.0, -1); copy a row only if its first id != 0.
static void TfPack(const uint64_t* eng_ids, const float* eng_w, const int32_t* eng_t,
                   int64_t rows, int32_t count, int64_t default_node, int64_t* out_ids,
                   float* out_w, int32_t* out_t) {
  for (int64_t i = 0; i < rows; ++i) {
    bool keep = eng_ids[i * count] != 0;
    for (int32_t j = 0; j < count; ++j) {
      int64_t o = i * count + j;
      out_ids[o] = keep ? (int64_t)eng_ids[o] : default_node;
      out_w[o] = keep ? eng_w[o] : 0.0f;
      out_t[o] = keep ? eng_t[o] : -1;
    }
  }
}

// tf_euler.sample_neighbor (tf_euler/kernels/sample_neighbor_op.cc:54-129)
void ref_op_sample_neighbor(const int64_t* nodes, int64_t n, const int32_t* etypes,
                            int32_t K, int32_t count, int64_t default_node,
                            int64_t* out_ids, float* out_w, int32_t* out_t) {
  std::vector<int> et(etypes, etypes + K);
  std::vector<uint64_t> e_ids(n * count);
  std::vector<float> e_w(n * count);
  std::vector<int32_t> e_t(n * count);
  EngineSampleNB(reinterpret_cast<const uint64_t*>(nodes), n, et, count, e_ids.data(),
                 e_w.data(), e_t.data());
  TfPack(e_ids.data(), e_w.data(), e_t.data(), n, count, default_node, out_ids, out_w, out_t);
}

// tf_euler.sample_fanout (tf_euler/kernels/sample_fanout_op.cc:36-43,116-140):
// hop i+1 seeds = hop i ENGINE ids (0 placeholders propagate).
// etypes is [L,K]; out_* are arrays of L pointers, hop i sized n*prod(counts[0..i]).
void ref_op_sample_fanout(const int64_t* nodes, int64_t n, const int32_t* etypes,
                          int32_t K, const int32_t* counts, int32_t L,
                          int64_t default_node, int64_t** out_ids, float** out_w,
                          int32_t** out_t) {
  std::vector<uint64_t> seeds(reinterpret_cast<const uint64_t*>(nodes),
                              reinterpret_cast<const uint64_t*>(nodes) + n);
  for (int32_t l = 0; l < L; ++l) {
    std::vector<int> et(etypes + l * K, etypes + (l + 1) * K);
    int64_t rows = seeds.size();
    int32_t c = counts[l];
    std::vector<uint64_t> e_ids(rows * c);
    std::vector<float> e_w(rows * c);
    std::vector<int32_t> e_t(rows * c);
    EngineSampleNB(seeds.data(), rows, et, c, e_ids.data(), e_w.data(), e_t.data());
    TfPack(e_ids.data(), e_w.data(), e_t.data(), rows, c, default_node, out_ids[l],
           out_w[l], out_t[l]);
    seeds.swap(e_ids);
  }
}

// node2vec step weights (tf_euler/kernels/random_walk_op.cc:140-168)
static void BuildWeights(const std::vector<int64_t>& pn, const std::vector<int64_t>& cn,
                         int64_t parent_id, float p, float q, std::vector<float>* w) {
  size_t j = 0, k = 0;
  while (j < cn.size() && k < pn.size()) {
    if (cn[j] < pn[k]) {
      if (cn[j] != parent_id) w->at(j) /= q; else w->at(j) /= p;
      ++j;
    } else if (cn[j] == pn[k]) {
      ++k; ++j;
    } else {
      ++k;
    }
  }
  while (j < cn.size()) {
    if (cn[j] != parent_id) w->at(j) /= q; else w->at(j) /= p;
    ++j;
  }
}

// tf_euler.random_walk (tf_euler/kernels/random_walk_op.cc:83-138,207-289).
// etypes is [L,K].  Neighbor lists come from euler::GetFullNeighbor; the
// engine's unique/gather wrap of API_GET_NB_NODE consumes no randomness.
void ref_op_random_walk(const int64_t* nodes, int64_t n, const int32_t* etypes, int32_t K,
                        int32_t L, float p, float q, int64_t default_node, int64_t* out) {
  for (int64_t i = 0; i < n; ++i) out[i * (L + 1)] = nodes[i];
  const float kEps = 1.0e-6;
  if (fabs(p - 1.0) <= kEps && fabs(q - 1.0) <= kEps) {
    // TraditionalRandomWalk: L chained sampleNB(count=1) hops (:207-247)
    std::vector<uint64_t> seeds(reinterpret_cast<const uint64_t*>(nodes),
                                reinterpret_cast<const uint64_t*>(nodes) + n);
    std::vector<uint64_t> e_ids(n);
    std::vector<float> e_w(n);
    std::vector<int32_t> e_t(n);
    for (int32_t l = 0; l < L; ++l) {
      std::vector<int> et(etypes + l * K, etypes + (l + 1) * K);
      EngineSampleNB(seeds.data(), n, et, 1, e_ids.data(), e_w.data(), e_t.data());
      for (int64_t i = 0; i < n; ++i)
        out[i * (L + 1) + l + 1] = e_ids[i] == 0 ? default_node : (int64_t)e_ids[i];
      seeds = e_ids;
    }
    return;
  }
  std::vector<std::vector<int64_t>> parent_neighbors(n);
  std::vector<int64_t> parent_ids(nodes, nodes + n);
  std::vector<int64_t> cur(nodes, nodes + n);
  for (int32_t step = 0; step < L; ++step) {
    std::vector<int> et(etypes + step * K, etypes + (step + 1) * K);
    euler::NodeIdVec q_ids(cur.begin(), cur.end());
    auto res = euler::GetFullNeighbor(q_ids, et);
    std::vector<std::vector<int64_t>> neighbors(n);
    std::vector<int64_t> next(n);
    for (int64_t i = 0; i < n; ++i) {
      std::vector<float> w;
      neighbors[i].reserve(res[i].size());
      w.reserve(res[i].size());
      for (auto& iw : res[i]) {
        neighbors[i].emplace_back((int64_t)std::get<0>(iw));
        w.emplace_back(std::get<1>(iw));
      }
      int64_t sample_id = default_node;
      if (!neighbors[i].empty()) {
        BuildWeights(parent_neighbors[i], neighbors[i], parent_ids[i], p, q, &w);
        euler::common::CompactWeightedCollection<int64_t> sampler;
        sampler.Init(neighbors[i], w);
        sample_id = sampler.Sample().first;
      }
      out[i * (L + 1) + step + 1] = sample_id;
      next[i] = sample_id;
    }
    parent_neighbors.swap(neighbors);
    parent_ids = cur;
    cur = next;
  }
}

// ------------------------------------------------------------- CPU baseline
// Timed loops for bench.py's reference arm.  Each worker thread runs `iters`
// independent batches (its own thread_local engine, like the reference's
// 8-thread client pool, euler/client/query_proxy.cc:205-210).  Seeds for batch
// b of thread t are seeds[((t*iters + b) % n_batches) * B ...].
// Returns wall seconds; *edges gets the number of sampled slots delivered.
double ref_bench_fanout(const int64_t* seeds, int64_t n_batches, int64_t B,
                        const int32_t* etypes, int32_t K, const int32_t* counts, int32_t L,
                        int32_t n_threads, int32_t iters, int32_t with_dedup,
                        int64_t* edges) {
  std::atomic<int64_t> total(0);
  auto worker = [&](int t) {
    ref_seed(12345 + t);
    int64_t local = 0;
    std::vector<std::vector<int64_t>> o_ids(L);
    std::vector<std::vector<float>> o_w(L);
    std::vector<std::vector<int32_t>> o_t(L);
    std::vector<int64_t*> p_ids(L);
    std::vector<float*> p_w(L);
    std::vector<int32_t*> p_t(L);
    int64_t rows = B;
    for (int l = 0; l < L; ++l) {
      rows *= counts[l];
      o_ids[l].resize(rows); o_w[l].resize(rows); o_t[l].resize(rows);
      p_ids[l] = o_ids[l].data(); p_w[l] = o_w[l].data(); p_t[l] = o_t[l].data();
      local += rows;
    }
    int64_t per_batch = local;
    local = 0;
    for (int b = 0; b < iters; ++b) {
      const int64_t* s = seeds + (((int64_t)t * iters + b) % n_batches) * B;
      if (with_dedup) {
        ref_op_sample_fanout(s, B, etypes, K, counts, L, -1, p_ids.data(), p_w.data(), p_t.data());
      } else {
        // raw api.cc chaining without the engine's unique/gather
        std::vector<uint64_t> cur(reinterpret_cast<const uint64_t*>(s),
                                  reinterpret_cast<const uint64_t*>(s) + B);
        for (int l = 0; l < L; ++l) {
          std::vector<int> et(etypes + l * K, etypes + (l + 1) * K);
          auto res = euler::SampleNeighbor(cur, et, counts[l]);
          std::vector<uint64_t> nxt;
          nxt.reserve(cur.size() * counts[l]);
          for (auto& row : res) {
            if (row.empty()) nxt.insert(nxt.end(), counts[l], 0);
            else for (auto& iw : row) nxt.push_back(std::get<0>(iw));
          }
          cur.swap(nxt);
        }
      }
      local += per_batch;
    }
    total += local;
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  *edges = total.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

// One full minibatch-construction step per batch, the way the reference's ops compose it
// (tf_euler/python/utils/encoders.py:475-491): sample_fanout, get_dense_feature for every hop, then
// the neighbor mean of each hop (scatter_add / (scatter_add(ones) + 1e-7), mp_ops.py:65-69 over
// scatter_op.cc:44-55).  Returns wall seconds; *edges = sampled slots delivered.
double ref_bench_step(const int64_t* seeds, int64_t n_batches, int64_t B, const int32_t* etypes,
                      int32_t K, const int32_t* counts, int32_t L, int32_t dim, int32_t n_threads,
                      int32_t iters, int64_t* edges) {
  std::atomic<int64_t> total(0);
  auto worker = [&](int t) {
    ref_seed(12345 + t);
    std::vector<std::vector<int64_t>> o_ids(L);
    std::vector<std::vector<float>> o_w(L);
    std::vector<std::vector<int32_t>> o_t(L);
    std::vector<int64_t*> p_ids(L);
    std::vector<float*> p_w(L);
    std::vector<int32_t*> p_t(L);
    std::vector<std::vector<float>> feat(L + 1), agg(L);
    std::vector<int64_t> rows(L + 1);
    rows[0] = B;
    int64_t per_batch = 0;
    for (int l = 0; l < L; ++l) {
      rows[l + 1] = rows[l] * counts[l];
      per_batch += rows[l + 1];
      o_ids[l].resize(rows[l + 1]); o_w[l].resize(rows[l + 1]); o_t[l].resize(rows[l + 1]);
      p_ids[l] = o_ids[l].data(); p_w[l] = o_w[l].data(); p_t[l] = o_t[l].data();
      agg[l].resize(rows[l] * (int64_t)dim);
    }
    for (int l = 0; l <= L; ++l) feat[l].resize(rows[l] * (int64_t)dim);
    std::vector<int32_t> len(rows[L]);
    std::vector<float> cnt;
    int64_t local = 0;
    for (int b = 0; b < iters; ++b) {
      const int64_t* s = seeds + (((int64_t)t * iters + b) % n_batches) * B;
      ref_op_sample_fanout(s, B, etypes, K, counts, L, -1, p_ids.data(), p_w.data(), p_t.data());
      for (int l = 0; l <= L; ++l) {
        const int64_t* ids = l == 0 ? s : p_ids[l - 1];
        ref_get_dense_feature(reinterpret_cast<const uint64_t*>(ids), rows[l], 0, dim, feat[l].data(), len.data());
      }
      for (int l = 0; l < L; ++l) {
        float* out = agg[l].data();
        std::fill_n(out, rows[l] * (int64_t)dim, 0);
        cnt.assign(rows[l], 0.f);
        const float* upd = feat[l + 1].data();
        for (int64_t i = 0; i < rows[l + 1]; ++i) {
          int64_t r = i / counts[l];
          for (int j = 0; j < dim; ++j) out[r * dim + j] += upd[i * dim + j];
          cnt[r] += 1.0f;
        }
        for (int64_t r = 0; r < rows[l]; ++r) {
          float c = cnt[r] + 1e-7f;
          for (int j = 0; j < dim; ++j) out[r * dim + j] = out[r * dim + j] / c;
        }
      }
      local += per_batch;
    }
    total += local;
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  *edges = total.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

// Dense feature fetch of `rows` ids through euler::GetNodeFloat32Feature,
// n_threads workers each doing `iters` passes.  Returns wall seconds.
double ref_bench_feature(const int64_t* ids, int64_t rows, int32_t dim, int32_t n_threads,
                         int32_t iters) {
  auto worker = [&](int) {
    std::vector<float> out(rows * (int64_t)dim);
    std::vector<int32_t> len(rows);
    for (int b = 0; b < iters; ++b)
      ref_get_dense_feature(reinterpret_cast<const uint64_t*>(ids), rows, 0, dim, out.data(), len.data());
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
