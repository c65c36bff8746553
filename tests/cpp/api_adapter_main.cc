// Exercises include/euler_b200_api.hpp the way a C++ caller of the reference's euler/core/api/api.h would
// (tests/test_gpu_parity.py::test_cpp_api_adapter compiles and runs it on the GPU box and checks the printed values
// against the oracle).  Output: one "key: v v v" line per query.
#include <cstdio>
#include <string>
#include <vector>

#include "euler_b200_api.hpp"

static void print_pairs(const char* key, const euler::IdWeightPairVec& v) {
  for (size_t i = 0; i < v.size(); ++i) {
    std::printf("%s[%zu]:", key, i);
    for (const auto& p : v[i]) std::printf(" %llu %.9g %d", (unsigned long long)std::get<0>(p), (double)std::get<1>(p), std::get<2>(p));
    std::printf("\n");
  }
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s <euler data dir> [count]\n", argv[0]); return 2; }
  const int count = argc > 2 ? std::atoi(argv[2]) : 5;
  if (!euler::InitGraph(argv[1])) { std::fprintf(stderr, "InitGraph failed: %s\n", eu_last_error()); return 1; }
  eu_ctx_seed(eu_default_ctx(), 4242);

  const euler::NodeIdVec nodes = {1, 2, 3, 4, 5, 6, 99};
  std::printf("node_type:");
  for (int32_t t : euler::GetNodeType(nodes)) std::printf(" %d", t);
  std::printf("\n");

  print_pairs("full", euler::GetFullNeighbor({1, 2, 99, 6}, {0, 1}));
  print_pairs("sample", euler::SampleNeighbor({1, 2, 3, 99, 1, 6}, {0, 1}, count));
  print_pairs("sample2", euler::SampleNeighbor({4, 5}, {1}, count));

  const euler::FloatFeatureVec f = euler::GetNodeFloat32Feature({1, 99, 3}, {0, 1, 7});
  for (size_t i = 0; i < f.size(); ++i)
    for (size_t k = 0; k < f[i].size(); ++k) {
      std::printf("feat[%zu][%zu]:", i, k);
      for (float x : f[i][k]) std::printf(" %.9g", (double)x);
      std::printf("\n");
    }

  std::printf("sample_node:");
  for (euler::NodeId id : euler::SampleNode({0, 1}, 8)) std::printf(" %llu", (unsigned long long)id);
  std::printf("\n");

  int t0 = -7, t1 = -7;
  const bool ok0 = euler::GetEdgeType(std::string(""), &t0);
  const bool ok1 = euler::GetNodeType(std::string("no such type"), &t1);
  std::printf("names: %d %d %d %d\n", (int)ok0, t0, (int)ok1, t1);

  // the op-level variant (ID_UNIQUE first: duplicates share a row)
  print_pairs("unique", euler::SampleNeighborUnique({1, 2, 3, 99, 1, 6}, {0, 1}, count));

  // euler::Graph / euler::Node (graph.h:41-93, node.h:63-110)
  euler::Graph& gr = euler::Graph::Instance();
  euler::Node* n1 = gr.GetNodeByID(1);
  euler::Node* n99 = gr.GetNodeByID(99);
  std::printf("graph_node: %d %llu %d %.9g %d\n", n1 != nullptr, n1 ? (unsigned long long)n1->GetID() : 0ull, n1 ? n1->GetType() : 0,
              n1 ? (double)n1->GetWeight() : 0.0, n99 == nullptr);
  if (n1) {
    euler::IdWeightPairVec one(1);
    one[0] = n1->SampleNeighbor({0, 1}, count);
    print_pairs("node_sample", one);
    one[0] = n1->GetFullNeighbor({0, 1});
    print_pairs("node_full", one);
    one[0] = n1->GetSortedFullNeighbor({0, 1});
    print_pairs("node_sorted", one);
    one[0] = n1->GetTopKNeighbor({0, 1}, 2);
    print_pairs("node_topk", one);
  }
  std::printf("graph_sample_node:");
  for (euler::NodeId id : gr.SampleNode(0, 6)) std::printf(" %llu", (unsigned long long)id);
  std::printf("\n");
  const euler::Status st = gr.Init(0, 1, "node", "/no/such/dir", "node");
  std::printf("graph_init_bad: %d\n", (int)st.ok());

  try {
    euler::SampleEdge({0}, 1);
    std::printf("out_of_scope: no throw\n");
  } catch (const std::runtime_error& e) {
    std::printf("out_of_scope: throws\n");
  }
  return 0;
}
