// euler_b200_api.hpp -- header-only C++ adapter: the reference's `euler/core/api/api.h` surface (namespace euler,
// same function names, argument and result types) over the euler_b200 C ABI, for the functions of the hot path.
// A C++ caller of the reference (tf_euler's kernels, euler/core/kernels/*.cc, user code) that includes this header
// instead of euler/core/api/api.h and links -leuler_b200 keeps compiling.
//
//   reference (euler/core/api/api.h)                      here
//   ----------------------------------------------------  ---------------------------------------------------------
//   NodeIdVec SampleNode(node_types, count)        :43    eu_sample_node_host
//   TypeVec GetNodeType(node_ids)                  :47    eu_get_node_type_host
//   FloatFeatureVec GetNodeFloat32Feature(ids,fids):50    eu_get_dense_feature_host, one call per slot
//   FloatFeatureVec GetNodeFloat32Feature(ids,names):57   names -> slots (eu_graph_dense_feature_id), then the above
//   IdWeightPairVec GetFullNeighbor(ids, etypes)   :78    eu_get_full_neighbor_host
//   IdWeightPairVec SampleNeighbor(ids, etypes, n) :80    eu_sample_neighbor_raw_host (api.cc:223-236: every occurrence of an id
//                                                         draws independently, in order -- results identical to the reference's
//                                                         serial loop under the same seed)
//   bool GetNodeType / GetEdgeType (names)         :83-89 eu_graph_node_type_id / eu_graph_edge_type_id
//   class Graph (euler/core/graph/graph.h:41-93)          euler::Graph: Instance(), Init(...), SampleNode(type|types, count),
//                                                         GetNodeByID(id) -> Node* proxy (nullptr when absent)
//   class Node  (euler/core/graph/node.h:63-110)          euler::Node: GetID/GetType/GetWeight, SampleNeighbor, GetFullNeighbor,
//                                                         GetSortedFullNeighbor, GetTopKNeighbor
//   graph start-up (Graph::Init, graph.h:53-60)           euler::InitGraph(data_path, shard_index, shard_number, device) / Graph::Init
//   SampleEdge, EdgeExist, edge / uint64 / binary features: not on the path -> std::runtime_error
//
//   euler::SampleNeighborUnique(ids, etypes, n)            the OP semantics every tf_euler op observes (engine rule ID_UNIQUE,
//                                                         euler/parser/compiler.cc:76-90): duplicate ids share one sampled row
//
// One behaviour differs from api.cc and is deliberate: results are deterministic under eu_ctx_seed
// (one serial minstd_rand0 stream per context), where api.cc's OpenMP build interleaves thread-local engines.
#ifndef EULER_B200_API_HPP_
#define EULER_B200_API_HPP_

#include <stdint.h>

#include <algorithm>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "euler_b200.h"

namespace euler {

using NodeId = uint64_t;                                   // euler/common/data_types.h:42
using EdgeId = std::tuple<NodeId, NodeId, int32_t>;        // :44
using IdWeightPair = std::tuple<NodeId, float, int32_t>;   // :46
typedef std::vector<std::vector<IdWeightPair>> IdWeightPairVec;
typedef std::vector<NodeId> NodeIdVec;
typedef std::vector<EdgeId> EdgeIdVec;
typedef std::vector<int32_t> TypeVec;
typedef std::vector<std::vector<std::vector<float>>> FloatFeatureVec;
typedef std::vector<std::vector<std::vector<uint64_t>>> UInt64FeatureVec;
typedef std::vector<std::vector<std::string>> BinaryFatureVec;

namespace b200_detail {
inline eu_ctx* ctx() {
  eu_ctx* c = eu_default_ctx();
  if (!c) throw std::runtime_error("euler_b200: no graph loaded (call euler::InitGraph or InitQueryProxy first)");
  return c;
}
inline void check(int rc) {
  if (rc != EU_OK) throw std::runtime_error(std::string("euler_b200: ") + eu_last_error());
}
inline const int64_t* i64(const NodeIdVec& v) { return reinterpret_cast<const int64_t*>(v.data()); }
[[noreturn]] inline void out_of_scope(const char* what) {
  throw std::runtime_error(std::string("euler_b200: ") + what + " is not part of the accelerated path");
}
}  // namespace b200_detail

// Graph::Init(shard_index, shard_number, "all"/"node", data_path, ...) (euler/core/graph/graph.h:53-60, graph.cc:90-98):
// loads the Euler 2.0 binary partitions of this shard into HBM and makes it the process default graph.
inline bool InitGraph(const std::string& data_path, int shard_index = 0, int shard_number = 1, int device = 0,
                      uint64_t seed = 1) {
  eu_graph* g = nullptr;
  if (eu_graph_load(data_path.c_str(), shard_index, shard_number, device, &g) != EU_OK) return false;
  return eu_set_default_graph(g, EU_RNG_MINSTD, seed) == EU_OK;
}

inline NodeIdVec SampleNode(const std::vector<int>& node_types, int count) {
  NodeIdVec out(count > 0 ? count : 0);
  std::vector<int32_t> t(node_types.begin(), node_types.end());
  b200_detail::check(eu_sample_node_host(b200_detail::ctx(), count, t.data(), (int32_t)t.size(), reinterpret_cast<int64_t*>(out.data())));
  return out;
}

inline TypeVec GetNodeType(const NodeIdVec& node_ids) {
  TypeVec out(node_ids.size());
  b200_detail::check(eu_get_node_type_host(b200_detail::ctx(), b200_detail::i64(node_ids), (int64_t)node_ids.size(), out.data()));
  return out;
}

// features[i][k] = the values of slot fids[k] of node i; a node that is not in the graph gets fids.size() empty vectors
// (api.cc:63-78), an unknown slot an empty vector.
inline FloatFeatureVec GetNodeFloat32Feature(const NodeIdVec& node_ids, const std::vector<int>& fids) {
  const size_t n = node_ids.size();
  FloatFeatureVec out(n, std::vector<std::vector<float>>(fids.size()));
  if (n == 0 || fids.empty()) return out;
  const TypeVec types = GetNodeType(node_ids);
  std::vector<float> buf;
  for (size_t k = 0; k < fids.size(); ++k) {
    const int32_t dim = eu_graph_dense_feature_dim(eu_default_graph(), fids[k]);
    if (dim <= 0) continue;
    buf.resize(n * (size_t)dim);
    b200_detail::check(eu_get_dense_feature_host(b200_detail::ctx(), b200_detail::i64(node_ids), (int64_t)n, fids[k], dim, buf.data()));
    for (size_t i = 0; i < n; ++i)
      if (types[i] != std::numeric_limits<int32_t>::lowest()) out[i][k].assign(buf.begin() + i * dim, buf.begin() + (i + 1) * dim);
  }
  return out;
}

inline FloatFeatureVec GetNodeFloat32Feature(const NodeIdVec& node_ids, const std::vector<std::string*>& ft_names) {
  std::vector<int> fids;
  for (const std::string* s : ft_names) fids.push_back(eu_graph_dense_feature_id(eu_default_graph(), s->c_str()));
  return GetNodeFloat32Feature(node_ids, fids);
}

inline IdWeightPairVec GetFullNeighbor(const NodeIdVec& node_ids, const std::vector<int>& edge_types) {
  const int64_t n = (int64_t)node_ids.size();
  std::vector<int32_t> et(edge_types.begin(), edge_types.end());
  std::vector<int64_t> ptr(n + 1, 0);
  int64_t total = 0;
  b200_detail::check(eu_get_full_neighbor_host(b200_detail::ctx(), b200_detail::i64(node_ids), n, et.data(), (int32_t)et.size(), 0,
                                               ptr.data(), nullptr, nullptr, nullptr, &total));
  std::vector<int64_t> ids(total > 0 ? total : 1);
  std::vector<float> w(ids.size());
  std::vector<int32_t> t(ids.size());
  if (total > 0)
    b200_detail::check(eu_get_full_neighbor_host(b200_detail::ctx(), b200_detail::i64(node_ids), n, et.data(), (int32_t)et.size(), total,
                                                 ptr.data(), ids.data(), w.data(), t.data(), &total));
  IdWeightPairVec out(n);
  for (int64_t i = 0; i < n; ++i)
    for (int64_t k = ptr[i]; k < ptr[i + 1]; ++k) out[i].emplace_back((NodeId)ids[k], w[k], t[k]);
  return out;
}

// neighbor[i] has `count` entries, or none when node i is absent / has no edge of the requested types (node.cc:98-161).
// api.cc:223-236: one Node::SampleNeighbor per element, in order; a repeated id draws again.
inline IdWeightPairVec SampleNeighbor(const NodeIdVec& node_ids, const std::vector<int>& edge_types, int count) {
  const size_t n = node_ids.size();
  IdWeightPairVec out(n);
  if (n == 0 || count <= 0) return out;
  std::vector<int32_t> et(edge_types.begin(), edge_types.end());
  std::vector<int64_t> ids(n * (size_t)count);
  std::vector<float> w(ids.size());
  std::vector<int32_t> t(ids.size());
  b200_detail::check(eu_sample_neighbor_raw_host(b200_detail::ctx(), b200_detail::i64(node_ids), (int64_t)n, et.data(), (int32_t)et.size(), count,
                                                 ids.data(), w.data(), t.data()));
  for (size_t i = 0; i < n; ++i)
    if (ids[i * count] != 0)   // engine form: a row whose first id is 0 (DEFAULT_UINT64) is an empty result
      for (int j = 0; j < count; ++j) out[i].emplace_back((NodeId)ids[i * count + j], w[i * count + j], t[i * count + j]);
  return out;
}

// The same call with the semantics every tf_euler OP observes: the engine uniquifies the ids first (ID_UNIQUE,
// euler/parser/compiler.cc:76-90), so duplicate ids in one call share one sampled row and consume one set of uniforms.
inline IdWeightPairVec SampleNeighborUnique(const NodeIdVec& node_ids, const std::vector<int>& edge_types, int count) {
  const size_t n = node_ids.size();
  IdWeightPairVec out(n);
  if (n == 0 || count <= 0) return out;
  std::vector<int32_t> et(edge_types.begin(), edge_types.end());
  std::vector<int64_t> ids(n * (size_t)count);
  std::vector<float> w(ids.size());
  std::vector<int32_t> t(ids.size());
  // default_node = 0 keeps the engine form: a row whose first id is 0 (DEFAULT_UINT64) is an empty result
  b200_detail::check(eu_sample_neighbor_host(b200_detail::ctx(), b200_detail::i64(node_ids), (int64_t)n, et.data(), (int32_t)et.size(), count,
                                             0, ids.data(), w.data(), t.data()));
  for (size_t i = 0; i < n; ++i)
    if (ids[i * count] != 0)
      for (int j = 0; j < count; ++j) out[i].emplace_back((NodeId)ids[i * count + j], w[i * count + j], t[i * count + j]);
  return out;
}

inline bool GetNodeType(const std::string& node_type, int* type_id) {
  if (node_type.empty()) { *type_id = -1; return true; }   // api.cc:264-268
  *type_id = eu_graph_node_type_id(eu_default_graph(), node_type.c_str());
  return *type_id >= 0;
}
inline bool GetEdgeType(const std::string& edge_type, int* type_id) {
  if (edge_type.empty()) { *type_id = -1; return true; }
  *type_id = eu_graph_edge_type_id(eu_default_graph(), edge_type.c_str());
  return *type_id >= 0;
}
inline bool GetNodeType(const std::vector<std::string*> node_types, std::vector<int>* type_ids) {
  type_ids->resize(node_types.size());
  for (size_t i = 0; i < node_types.size(); ++i)
    if (!GetNodeType(*node_types[i], &type_ids->at(i)) || type_ids->at(i) < 0) { type_ids->clear(); return false; }
  return true;
}
inline bool GetEdgeType(const std::vector<std::string*> edge_types, std::vector<int>* type_ids) {
  type_ids->resize(edge_types.size());
  for (size_t i = 0; i < edge_types.size(); ++i)
    if (!GetEdgeType(*edge_types[i], &type_ids->at(i)) || type_ids->at(i) < 0) { type_ids->clear(); return false; }
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// euler::Graph / euler::Node (euler/core/graph/graph.h:41-93, node.h:63-110) for C++ code written against the classes
// rather than the free functions.  The graph lives in HBM; a Node is a light proxy (id, type, weight) whose methods are
// one-row calls into the same kernels.  Status mirrors euler/common/status.h's ok()/error_message().
class Status {
 public:
  Status() : ok_(true) {}
  explicit Status(const std::string& msg) : ok_(false), msg_(msg) {}
  static Status OK() { return Status(); }
  bool ok() const { return ok_; }
  const std::string& error_message() const { return msg_; }
 private:
  bool ok_;
  std::string msg_;
};

namespace common {
typedef uint64_t NodeID;
typedef std::tuple<NodeID, float, int32_t> IDWeightPair;
}  // namespace common

class Node {
 public:
  Node(common::NodeID id, float weight, int32_t type) : id_(id), weight_(weight), type_(type) {}
  common::NodeID GetID() const { return id_; }
  int32_t GetType() const { return type_; }
  float GetWeight() const { return weight_; }
  // node.cc:98-161 (count entries, or none)
  std::vector<common::IDWeightPair> SampleNeighbor(const std::vector<int32_t>& edge_types, int32_t count) const {
    return SampleNeighbor_(NodeIdVec{id_}, edge_types, count);
  }
  // node.cc:176-198
  std::vector<common::IDWeightPair> GetFullNeighbor(const std::vector<int32_t>& edge_types) const {
    return euler::GetFullNeighbor(NodeIdVec{id_}, std::vector<int>(edge_types.begin(), edge_types.end()))[0];
  }
  // node.cc:210-262: ordered by neighbor id
  std::vector<common::IDWeightPair> GetSortedFullNeighbor(const std::vector<int32_t>& edge_types) const {
    std::vector<common::IDWeightPair> v = GetFullNeighbor(edge_types);
    std::stable_sort(v.begin(), v.end(), [](const common::IDWeightPair& a, const common::IDWeightPair& b) { return std::get<0>(a) < std::get<0>(b); });
    return v;
  }
  // node.cc:264-316: the k heaviest, heaviest first
  std::vector<common::IDWeightPair> GetTopKNeighbor(const std::vector<int32_t>& edge_types, int32_t k) const {
    std::vector<common::IDWeightPair> v = GetFullNeighbor(edge_types);
    std::stable_sort(v.begin(), v.end(), [](const common::IDWeightPair& a, const common::IDWeightPair& b) { return std::get<1>(a) > std::get<1>(b); });
    if (k >= 0 && (size_t)k < v.size()) v.resize(k);
    if (k <= 0) v.clear();
    return v;
  }
 private:
  static std::vector<common::IDWeightPair> SampleNeighbor_(const NodeIdVec& one, const std::vector<int32_t>& edge_types, int32_t count) {
    return euler::SampleNeighbor(one, std::vector<int>(edge_types.begin(), edge_types.end()), count)[0];
  }
  common::NodeID id_;
  float weight_;
  int32_t type_;
};

class Graph {
 public:
  static Graph& Instance() {   // graph.h:64-67
    static Graph instance;
    return instance;
  }
  // graph.h:53-56.  sampler_type / load_data_type are accepted for source compatibility: the node sampler is built on first
  // use and this path loads node data only.  `device` and `seed` are extensions with defaults.
  Status Init(int shard_index, int shard_number, const std::string& sampler_type, const std::string& data_path,
              const std::string& load_data_type, int device = 0, uint64_t seed = 1) {
    (void)sampler_type; (void)load_data_type;
    if (!InitGraph(data_path, shard_index, shard_number, device, seed)) return Status(std::string("Graph::Init: ") + eu_last_error());
    std::lock_guard<std::mutex> l(mu_);
    nodes_.clear();
    return Status::OK();
  }
  // graph.h:75-79, graph.cc:221-275
  std::vector<common::NodeID> SampleNode(int node_type, int count) const { return euler::SampleNode(std::vector<int>{node_type}, count); }
  std::vector<common::NodeID> SampleNode(const std::vector<int>& node_types, int count) const { return euler::SampleNode(node_types, count); }
  // graph.h:87-93: nullptr when the id is not a node.  The proxy stays valid until the next Init.
  Node* GetNodeByID(common::NodeID id) const {
    std::lock_guard<std::mutex> l(mu_);
    auto it = nodes_.find(id);
    if (it != nodes_.end()) return it->second.get();
    int32_t type = 0;
    float weight = 0.f;
    const int64_t sid = (int64_t)id;
    b200_detail::check(eu_get_node_type_host(b200_detail::ctx(), &sid, 1, &type));
    if (type == std::numeric_limits<int32_t>::lowest()) return nullptr;
    b200_detail::check(eu_get_node_weight_host(b200_detail::ctx(), &sid, 1, &weight));
    Node* n = new Node(id, weight, type);
    nodes_[id].reset(n);
    return n;
  }
 private:
  Graph() {}
  mutable std::mutex mu_;
  mutable std::unordered_map<common::NodeID, std::unique_ptr<Node>> nodes_;
};

// ---- declared by api.h, outside the accelerated path ("out of scope")
inline bool EdgeExist(const EdgeId&) { b200_detail::out_of_scope("EdgeExist"); }
inline EdgeIdVec SampleEdge(const std::vector<int>&, int) { b200_detail::out_of_scope("SampleEdge"); }
inline UInt64FeatureVec GetNodeUint64Feature(const NodeIdVec&, const std::vector<int>&) { b200_detail::out_of_scope("GetNodeUint64Feature"); }
inline BinaryFatureVec GetNodeBinaryFeature(const NodeIdVec&, const std::vector<int>&) { b200_detail::out_of_scope("GetNodeBinaryFeature"); }
inline FloatFeatureVec GetEdgeFloat32Feature(const EdgeIdVec&, const std::vector<int>&) { b200_detail::out_of_scope("GetEdgeFloat32Feature"); }
inline UInt64FeatureVec GetEdgeUint64Feature(const EdgeIdVec&, const std::vector<int>&) { b200_detail::out_of_scope("GetEdgeUint64Feature"); }
inline BinaryFatureVec GetEdgeBinaryFeature(const EdgeIdVec&, const std::vector<int>&) { b200_detail::out_of_scope("GetEdgeBinaryFeature"); }

}  // namespace euler

#endif  // EULER_B200_API_HPP_
